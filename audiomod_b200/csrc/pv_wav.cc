// File -> file batch front end (include/pvgpu.h: pvgpu_run_wav_files): what `audiomod-exe <effect> in.wav out.wav ...` does
// for one file, for thousands of files at once.
//
// Host side only (the device work is pvgpu_mbatch / pvgpu_batch).  Follows the reference CLI and its WAV reader / writer:
//   header walk        main/wavfile.cc:848-1016 (RIFF/WAVE, 'fmt ' truncated to 16 bytes, 'fact', unknown chunks skipped by
//                      their length without a pad byte, labels must be printable, 'data' ends the walk)
//   length             getNumSamples, wavfile.cc:1058-1063: data_len / byte_per_sample (fact_sample_len when format tag > 1)
//   sample decode      wavfile.cc:672-812: 8 bit (u8 / 128 - 1), 16, 24, 32 bit little endian, value * (1 / 2^(bits-1)) in double
//   block protocol     main/main.cc:149,471-509 (inside pvgpu_batch_plan)
//   output             always 16-bit PCM (main.cc:136), header of wavfile.cc:1135-1171 (RIFF + fmt(16) + fact(4) + data = 56
//                      bytes), samples (short)(int)clamp(x * 32768.f) truncating toward zero (wavfile.cc:1294-1306, 1508-1526)
// 16-bit files travel to the GPU as int16 rows and come back as int16 rows (the conversions above run on the device,
// pv_synth.cuh pcm_load / pcm_store); other widths are decoded to float32 on the host exactly like the reference's reader.
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <map>
#include <new>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/pvgpu.h"
#include "pv_internal.h"

namespace pvgpu {

struct WavInfo {
    int format_tag = 0, channels = 0, sample_rate = 0, byte_rate = 0, block_align = 0, bits = 0;
    uint32_t fact_samples = 0, data_len = 0;
    long data_pos = 0;
    bool have_fmt = false, have_data = false;
};

static bool printable_label(const char *l) {
    for (int i = 0; i < 4; ++i) if (l[i] < ' ' || l[i] > 'z') return false;
    return true;
}
static uint32_t le32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t le16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

// returns an empty string on success
static std::string parse_header(FILE *f, WavInfo &w) {
    unsigned char b[16];
    if (fread(b, 1, 12, f) != 12 || memcmp(b, "RIFF", 4) != 0 || memcmp(b + 8, "WAVE", 4) != 0) return "not a RIFF/WAVE file";
    for (;;) {
        char label[4];
        if (fread(label, 1, 4, f) != 4) return "no data chunk";
        if (!printable_label(label)) return "invalid chunk label";
        if (memcmp(label, "INFO", 4) == 0) continue;     // the reference does nothing for this label, not even read a length (:931-934)
        unsigned char lb[4];
        if (fread(lb, 1, 4, f) != 4) return "truncated chunk header";
        const uint32_t len = le32(lb);
        if (memcmp(label, "fmt ", 4) == 0) {
            const uint32_t take = len > 16 ? 16 : len;
            memset(b, 0, sizeof b);
            if (fread(b, 1, take, f) != take) return "truncated fmt chunk";
            w.format_tag = le16(b); w.channels = le16(b + 2); w.sample_rate = (int)le32(b + 4); w.byte_rate = (int)le32(b + 8);
            w.block_align = le16(b + 12); w.bits = le16(b + 14);
            w.have_fmt = true;
            if (len > 16 && fseek(f, (long)(len - 16), SEEK_CUR) != 0) return "truncated fmt chunk";
        } else if (memcmp(label, "fact", 4) == 0) {
            const uint32_t take = len > 4 ? 4 : len;
            memset(b, 0, sizeof b);
            if (fread(b, 1, take, f) != take) return "truncated fact chunk";
            w.fact_samples = le32(b);
            if (len > 4 && fseek(f, (long)(len - 4), SEEK_CUR) != 0) return "truncated fact chunk";
        } else if (memcmp(label, "data", 4) == 0) {
            w.data_len = len;
            w.data_pos = ftell(f);
            w.have_data = true;
            break;
        } else {
            if (fseek(f, (long)len, SEEK_CUR) != 0) return "truncated chunk";   // no pad byte, like the reference
        }
    }
    if (!w.have_fmt) return "no fmt chunk";
    return "";
}

struct Job {
    pvgpu_wav_job *user;
    WavInfo w;
    int64_t frames = 0;               // file_length of main.cc:134
    std::vector<unsigned char> raw;   // the data chunk
};

static void set_status(pvgpu_wav_job *j, int code, const std::string &msg) {
    j->status = code;
    snprintf(j->message, sizeof j->message, "%s", msg.c_str());
}

// run fn(i) for i in [0, n) on up to `threads` host threads
template <class Fn> static void parallel_for(int n, int threads, Fn fn) {
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    const int nt = n < threads ? (n < 1 ? 1 : n) : threads;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&]() { for (int i = next++; i < n; i = next++) fn(i); });
    for (auto &t : th) t.join();
}

static std::string write_wav16(const char *path, int sample_rate, int channels, const short *const *rows, int64_t frames) {
    FILE *f = fopen(path, "wb");
    if (!f) return std::string("cannot open ") + path + " for writing";
    const uint32_t bytes = (uint32_t)(frames * channels * 2);
    unsigned char h[56];
    auto p32 = [&](int o, uint32_t v) { h[o] = v & 255; h[o + 1] = (v >> 8) & 255; h[o + 2] = (v >> 16) & 255; h[o + 3] = (v >> 24) & 255; };
    auto p16 = [&](int o, uint32_t v) { h[o] = v & 255; h[o + 1] = (v >> 8) & 255; };
    memcpy(h, "RIFF", 4); p32(4, bytes + 56 - 12 + 4); memcpy(h + 8, "WAVE", 4);        // finishHeader, wavfile.cc:1174-1182
    memcpy(h + 12, "fmt ", 4); p32(16, 16); p16(20, 1); p16(22, (uint32_t)channels); p32(24, (uint32_t)sample_rate);
    p32(28, (uint32_t)(2 * channels * sample_rate)); p16(32, (uint32_t)(2 * channels)); p16(34, 16);
    memcpy(h + 36, "fact", 4); p32(40, 4); p32(44, bytes / (uint32_t)(2 * channels));
    memcpy(h + 48, "data", 4); p32(52, bytes);
    bool ok = fwrite(h, 1, 56, f) == 56;
    std::vector<short> inter((size_t)frames * channels);
    for (int c = 0; c < channels; ++c)
        for (int64_t i = 0; i < frames; ++i) inter[(size_t)i * channels + c] = rows[c][i];
    ok = ok && fwrite(inter.data(), 2, inter.size(), f) == inter.size();
    ok = fclose(f) == 0 && ok;
    return ok ? "" : std::string("short write to ") + path;
}

}  // namespace pvgpu

using namespace pvgpu;

static int run_wav_files_impl(const pvgpu_config *cfg, pvgpu_wav_job *jobs, int n_jobs, const int *devices, int n_dev);

extern "C" int pvgpu_run_wav_files(const pvgpu_config *cfg, pvgpu_wav_job *jobs, int n_jobs, const int *devices, int n_dev) {
    try {   // no exception may cross the C ABI (the reference aborts on allocation failure, memallocators.h:91; here: an error code)
        return run_wav_files_impl(cfg, jobs, n_jobs, devices, n_dev);
    } catch (const std::bad_alloc &) {
        return fail(PVGPU_ENOMEM, "out of host memory while reading the WAV files");
    } catch (const std::exception &e) {
        return fail(PVGPU_ESTATE, "unexpected failure: %s", e.what());
    }
}

static int run_wav_files_impl(const pvgpu_config *cfg, pvgpu_wav_job *jobs, int n_jobs, const int *devices, int n_dev) {
    if (!cfg || !jobs || n_jobs < 0) return fail(PVGPU_EINVAL, "bad argument");
    if (pvgpu_device_count() < 1) return fail(PVGPU_ECUDA, "no CUDA device: the phase vocoder has no CPU fallback");
    std::vector<Job> js((size_t)n_jobs);
    const int io_threads = 16;
    // ---- read: header walk + the data chunk ----
    parallel_for(n_jobs, io_threads, [&](int i) {
        Job &j = js[i];
        j.user = &jobs[i];
        jobs[i].status = PVGPU_OK; jobs[i].message[0] = 0; jobs[i].frames_in = jobs[i].frames_out = 0;
        jobs[i].sample_rate = jobs[i].channels = jobs[i].bits = 0;
        FILE *f = jobs[i].in_path ? fopen(jobs[i].in_path, "rb") : nullptr;
        if (!f) { set_status(&jobs[i], PVGPU_EINVAL, "cannot open input file"); return; }
        std::string err = parse_header(f, j.w);
        const WavInfo &w = j.w;
        if (err.empty() && (w.bits != 8 && w.bits != 16 && w.bits != 24 && w.bits != 32)) err = "only 8/16/24/32 bit PCM is supported (wavfile.cc:684-691)";
        if (err.empty() && (w.channels < 1 || w.channels > 2)) err = "only mono and stereo files are supported (the reference's reader de-interleaves two channels)";
        if (err.empty() && w.format_tag != 1) err = "only format tag 1 (integer PCM) is supported";
        if (err.empty() && w.block_align == 0) err = "byte_per_sample is zero";
        if (err.empty()) {
            j.frames = (int64_t)(w.data_len / (uint32_t)(unsigned short)w.block_align);
            const size_t want = (size_t)j.frames * (size_t)w.channels * (size_t)(w.bits / 8);
            j.raw.resize(want);
            if (want && fread(j.raw.data(), 1, want, f) != want) err = "data chunk is shorter than its header says";
        }
        fclose(f);
        jobs[i].sample_rate = w.sample_rate; jobs[i].channels = w.channels; jobs[i].bits = w.bits; jobs[i].frames_in = j.frames;
        if (!err.empty()) set_status(&jobs[i], PVGPU_EINVAL, err);
    });
    // ---- group by (sample rate, channels, 16-bit or not): one configuration per batch ----
    std::map<std::tuple<int, int, bool>, std::vector<int>> groups;
    for (int i = 0; i < n_jobs; ++i)
        if (jobs[i].status == PVGPU_OK) groups[std::make_tuple(js[i].w.sample_rate, js[i].w.channels, js[i].w.bits == 16)].push_back(i);
    int first_error = PVGPU_OK;
    std::string first_msg;
    for (auto &kv : groups) {
        const int sr = std::get<0>(kv.first), C = std::get<1>(kv.first);
        const bool s16 = std::get<2>(kv.first);
        const std::vector<int> &mem = kv.second;
        const int S = (int)mem.size();
        auto group_fail = [&](int code, const std::string &msg) {
            for (int i : mem) set_status(&jobs[i], code, msg);
            if (first_error == PVGPU_OK) { first_error = code; first_msg = msg; }
        };
        pvgpu_config c = *cfg;
        c.sample_rate = sr; c.channels = C;
        std::vector<int64_t> n_in(S), n_out(S);
        int64_t longest = 0;
        for (int s = 0; s < S; ++s) { n_in[s] = js[mem[s]].frames; longest = n_in[s] > longest ? n_in[s] : longest; }
        pvgpu_mbatch *mb = nullptr;
        int rc = pvgpu_mbatch_create(&c, S, longest, devices, n_dev, &mb);
        if (rc == PVGPU_OK) rc = pvgpu_mbatch_plan(mb, n_in.data(), 0, n_out.data());
        if (rc != PVGPU_OK) { group_fail(rc, pvgpu_last_error()); if (mb) pvgpu_mbatch_destroy(mb); continue; }
        // planar rows in one page-locked slab per direction
        const size_t esz = s16 ? 2 : 4;
        std::vector<size_t> in_off((size_t)S * C), out_off((size_t)S * C);
        size_t in_bytes = 0, out_bytes = 0;
        for (int s = 0; s < S; ++s)
            for (int ch = 0; ch < C; ++ch) {
                in_off[(size_t)s * C + ch] = in_bytes; in_bytes += (((size_t)n_in[s] * esz) + 63) & ~(size_t)63;
                out_off[(size_t)s * C + ch] = out_bytes; out_bytes += (((size_t)n_out[s] * esz) + 63) & ~(size_t)63;
            }
        void *slab_in = nullptr, *slab_out = nullptr;
        const int dev0 = (devices && n_dev > 0) ? devices[0] : 0;
        rc = pvgpu_host_alloc(&slab_in, in_bytes ? in_bytes : 64, dev0, 0);
        if (rc == PVGPU_OK) rc = pvgpu_host_alloc(&slab_out, out_bytes ? out_bytes : 64, dev0, 0);
        if (rc != PVGPU_OK) { group_fail(rc, pvgpu_last_error()); pvgpu_host_free(slab_in); pvgpu_mbatch_destroy(mb); continue; }
        // ---- decode + de-interleave (wavfile.cc:672-812) ----
        parallel_for(S, io_threads, [&](int s) {
            const Job &j = js[mem[s]];
            const unsigned char *raw = j.raw.data();
            const int bps = j.w.bits / 8;
            for (int ch = 0; ch < C; ++ch) {
                char *dst = (char *)slab_in + in_off[(size_t)s * C + ch];
                for (int64_t i = 0; i < j.frames; ++i) {
                    const unsigned char *p = raw + ((size_t)i * C + ch) * bps;
                    if (s16) { ((short *)dst)[i] = (short)le16(p); continue; }
                    float v;
                    if (bps == 1) v = (float)(p[0] * (1.0 / 128.0) - 1.0);
                    else if (bps == 3) { int x = (int)(p[0] | (p[1] << 8) | (p[2] << 16)); if (x & 0x800000) x |= ~0xffffff; v = (float)(x * (1.0 / 8388608.0)); }
                    else v = (float)((int)le32(p) * (1.0 / 2147483648.0));
                    ((float *)dst)[i] = v;
                }
            }
        });
        for (int i : mem) std::vector<unsigned char>().swap(js[i].raw);
        std::vector<const void *> in_rows((size_t)S * C);
        std::vector<void *> out_rows((size_t)S * C);
        for (size_t r = 0; r < in_rows.size(); ++r) { in_rows[r] = (char *)slab_in + in_off[r]; out_rows[r] = (char *)slab_out + out_off[r]; }
        rc = pvgpu_mbatch_run_host(mb, in_rows.data(), out_rows.data(), s16 ? PVGPU_S16 : PVGPU_F32);
        if (rc != PVGPU_OK) group_fail(rc, pvgpu_last_error());
        else {
            // ---- 16-bit output files ----
            parallel_for(S, io_threads, [&](int s) {
                pvgpu_wav_job *u = js[mem[s]].user;
                const int64_t n = n_out[s];
                std::vector<short> conv;
                const short *rows[2] = {nullptr, nullptr};
                if (!s16) conv.resize((size_t)n * C);
                for (int ch = 0; ch < C; ++ch) {
                    const char *src = (const char *)slab_out + out_off[(size_t)s * C + ch];
                    if (s16) { rows[ch] = (const short *)src; continue; }
                    short *d = conv.data() + (size_t)ch * n;
                    for (int64_t i = 0; i < n; ++i) {   // saturate(), wavfile.cc:1294-1306
                        float x = ((const float *)src)[i] * 32768.0f;
                        x = x > 32767.0f ? 32767.0f : (x < -32768.0f ? -32768.0f : x);
                        d[i] = (short)(int)x;
                    }
                    rows[ch] = d;
                }
                const std::string err = u->out_path ? write_wav16(u->out_path, sr, C, rows, n) : std::string("no output path");
                u->frames_out = n;
                if (!err.empty()) set_status(u, PVGPU_EINVAL, err);
            });
        }
        pvgpu_host_free(slab_in);
        pvgpu_host_free(slab_out);
        pvgpu_mbatch_destroy(mb);
    }
    if (first_error != PVGPU_OK) return fail(first_error, "%s", first_msg.c_str());
    for (int i = 0; i < n_jobs; ++i)
        if (jobs[i].status != PVGPU_OK) return fail(jobs[i].status, "%s: %s", jobs[i].in_path ? jobs[i].in_path : "(null)", jobs[i].message);
    return PVGPU_OK;
}
