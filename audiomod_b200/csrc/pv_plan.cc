// Host-side plan: see pv_plan.h.  Plain C++; compile with -ffp-contract=off so that float
// expressions round exactly like the reference's default (SSE2, no FMA) build.
#include "pv_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace pvgpu {

// ---------------------------------------------------------------------------------------
// sizes: phasevocoder::phasevocoder (phasevocoder.cc:24-60) + Impl::calculateSizes
// (phasevocoderimpl.cc:169-263)
// ---------------------------------------------------------------------------------------
static int next_pow2(size_t v) {
    if (!(v & (v - 1))) return (int)v;
    int bits = 0;
    while (v) { ++bits; v >>= 1; }
    return 1 << bits;
}

static uint32_t gcd_u32(uint32_t a, uint32_t b) {
    while (b) { uint32_t t = b; b = a % b; a = t; }
    return a;
}

// Kaiser-8 window samples the Speex resampler interpolates (speex/resample.c:229-236).
static const double kKaiser8[36] = {
    0.99635258, 1.00000000, 0.99635258, 0.98548012, 0.96759014, 0.94302200, 0.91223751, 0.87580811, 0.83439927,
    0.78875245, 0.73966538, 0.68797126, 0.63451750, 0.58014482, 0.52566725, 0.47185369, 0.41941150, 0.36897272,
    0.32108304, 0.27619388, 0.23465776, 0.19672670, 0.16255380, 0.13219758, 0.10562887, 0.08273982, 0.06335451,
    0.04724088, 0.03412321, 0.02369490, 0.01563093, 0.00959968, 0.00527363, 0.00233883, 0.00050000, 0.00000000};

static double kaiser8_at(float x) {  // compute_func, resample.c:300-322 (window oversampling 32)
    const float y = x * 32;
    const int ind = (int)floor(y);
    const float fr = y - ind;
    double c[4];
    c[3] = -0.1666666667 * fr + 0.1666666667 * (fr * fr * fr);
    c[2] = fr + 0.5 * (fr * fr) - 0.5 * (fr * fr * fr);
    c[0] = -0.3333333333 * fr + 0.5 * (fr * fr) - 0.1666666667 * (fr * fr * fr);
    c[1] = 1.f - c[3] - c[2] - c[0];
    return c[0] * kKaiser8[ind] + c[1] * kKaiser8[ind + 1] + c[2] * kKaiser8[ind + 2] + c[3] * kKaiser8[ind + 3];
}

static float windowed_sinc(float cutoff, float x, int len) {  // sinc, resample.c:325-337
    const float xx = x * cutoff;
    if (fabsf(x) < 1e-6) return cutoff;
    if (fabsf(x) > .5 * len) return 0;
    return cutoff * sin(M_PI * xx) / (M_PI * xx) * kaiser8_at(fabs(2. * x / len));
}

// RS_Speex::setratio (resampler.cc:740-770) -> speex_resampler_set_rate_frac
// (resample.c:1117-1158) -> update_filter (resample.c:661-779), quality 4 = {64, 8, .921, .940, Kaiser8}.
static void build_resampler(ResamplerSpec &rs, float ratio) {
    const unsigned big = 272408136U;
    unsigned denom = 1, numer = 1;
    if (ratio < 1.f) {
        denom = big;
        numer = (unsigned)((double)big * (double)ratio);
    } else if (ratio > 1.f) {
        numer = big;
        denom = (unsigned)((double)big / (double)ratio);
    }
    rs.ratio = ratio;
    rs.num = denom;  // input rate
    rs.den = numer;  // output rate
    const uint32_t g = gcd_u32(rs.num, rs.den);
    rs.num /= g;
    rs.den /= g;
    rs.oversample = 8;
    rs.filt_len = 64;
    if (rs.num > rs.den) {
        rs.cutoff = 0.921f * rs.den / rs.num;
        rs.filt_len = (unsigned)ceil(rs.filt_len * ((double)rs.num / (double)rs.den));
        rs.filt_len &= ~0x3u;
        if (2 * rs.den < rs.num) rs.oversample >>= 1;
        if (4 * rs.den < rs.num) rs.oversample >>= 1;
        if (8 * rs.den < rs.num) rs.oversample >>= 1;
        if (16 * rs.den < rs.num) rs.oversample >>= 1;
        if (rs.oversample < 1) rs.oversample = 1;
    } else {
        rs.cutoff = 0.940f;
    }
    const int L = (int)rs.filt_len, ov = (int)rs.oversample;
    if (rs.den <= rs.oversample) {
        rs.direct = true;
        rs.table.resize((size_t)L * rs.den);
        for (uint32_t i = 0; i < rs.den; ++i)
            for (int j = 0; j < L; ++j)
                rs.table[(size_t)i * L + j] = windowed_sinc(rs.cutoff, ((j - L / 2 + 1) - ((float)i) / rs.den), L);
    } else {
        rs.direct = false;
        rs.table.resize((size_t)L * ov + 8);
        for (int k = -4; k < ov * L + 4; ++k)
            rs.table[k + 4] = windowed_sinc(rs.cutoff, (k / (float)ov - L / 2), L);
    }
    rs.int_adv = (int)(rs.num / rs.den);
    rs.frac_adv = (int)(rs.num % rs.den);
}

Derived derive(const Config &cfg) {
    Derived d;
    d.cfg = cfg;
    d.valid_mode = cfg.mode >= -1 && cfg.mode <= kFormantCepstral;
    float time_ratio = cfg.time_ratio;
    float pitch_scale = cfg.pitch_semitones != 0 ? (float)std::pow(2.0, cfg.pitch_semitones / 12) : 1.0f;
    d.gender = cfg.mode == kGender || cfg.mode == kGenderCepstral;
    d.formant = cfg.mode == kFormant || cfg.mode == kFormantCepstral;
    d.robotic = cfg.mode == kRobotic;
    d.whisper = cfg.mode == kWhisper;
    d.vocoder = cfg.mode == kVocRosen || cfg.mode == kVocChord;
    d.constant_mode = cfg.mode == kConstant;
    const size_t window = (size_t)next_pow2((size_t)(cfg.fftsize > 0 ? cfg.fftsize : 1));
    if (pitch_scale <= 0.0) pitch_scale = 1.0;
    if (time_ratio <= 0.0) time_ratio = 1.0;
    const float hs = time_ratio * pitch_scale;
    size_t in_hop, out_hop;
    if (cfg.hopsize > 0) {
        in_hop = (size_t)cfg.hopsize;
        out_hop = (size_t)(int)(floor(in_hop * hs));
    } else {
        float wir;
        if (hs < 1) {
            wir = pitch_scale < 1.0 ? 4.5f : 6.f;
            in_hop = (size_t)(int)(window / wir);
            out_hop = (size_t)(int)(in_hop * hs);
        } else {
            wir = hs == 1.0 ? 4.f : 8.f;
            out_hop = (size_t)(int)(window / wir);
            in_hop = (size_t)(int)(out_hop / hs);
        }
    }
    (void)out_hop;
    d.N = (int)window;
    d.H = d.N / 2 + 1;
    d.hop = (int)in_hop;
    d.pitch_scale = pitch_scale;
    d.hs = hs;
    d.outbuf_cap = hs > 1 ? (long)(size_t)(window * 16 * hs) : (long)(window * 16);
    if (d.outbuf_cap < 2L * d.N) d.outbuf_cap = 2L * d.N;  // channelinfo.cc:30-36
    d.int_ratio = fabsf(hs - floorf(hs)) <= 0.001;          // Impl::isIntRatio, impl.cc:149-157
    // frequency-axis warp factor: synthesiseSlice, phasevocoderprocess.cc:1006-1022
    d.freq_comp = 0.f;
    if (d.formant && pitch_scale != 1.0) d.freq_comp = pitch_scale;
    if (d.gender && pitch_scale != 1.0) d.freq_comp = pitch_scale > 1 ? (float)(0.85 * pitch_scale) : (float)(1.17 * pitch_scale);
    else if (d.gender) d.freq_comp = (float)0.8;
    if ((cfg.mode == kGenderCepstral || cfg.mode == kFormantCepstral) && pitch_scale != 1.0) {
        // the cepstral variants of maleToFemale / femaleToMale / formantPreserveSlice (:824-840 with the comments swapped)
        d.freq_comp = 0.f;
        d.cepstral = true;
        d.env_comp = cfg.mode == kFormantCepstral ? 1.f : (pitch_scale > 1 ? 0.85f : 1.17f);
    }
    d.fixed_gain = pitch_scale > 1 ? pitch_scale : 1 / pitch_scale;
    d.rs.active = pitch_scale != 1.0;
    if (d.rs.active) build_resampler(d.rs, (float)(1.0 / pitch_scale));
    return d;
}

// ---------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------
Tables make_tables(int N, int hop) {
    Tables t;
    const int nc = N / 2;
    // Hann: windowfunc<float>(Hanning, N), windowfunc.h:101-169
    t.window.assign(N, 1.0f);
    for (int i = 0; i < N; ++i) t.window[i] *= (0.50f - 0.50f * cos(2 * M_PI * i / N) + 0.0f * cos(4 * M_PI * i / N) - 0.0f * cos(6 * M_PI * i / N));
    float area = 0;
    for (int i = 0; i < N; ++i) area += t.window[i];
    area /= N;
    t.window_area = area;
    t.acc_scale = area * 1.5;
    // factorisation (kf_factor, kiss_fft.c:292-314): 4s first, then a 2; executed innermost first
    std::vector<int> fp, fm;
    {
        int rem = nc, q = 4;
        do {
            while (rem % q) q = 2;
            rem /= q;
            fp.push_back(q);
            fm.push_back(rem);
        } while (rem > 1);
    }
    const int nf = (int)fp.size();
    for (int i = 0; i < nf; ++i) { t.radix.push_back(fp[nf - 1 - i]); t.span.push_back(fm[nf - 1 - i]); }
    t.perm.resize(nc);
    for (int o = 0; o < nc; ++o) {  // kf_work recursion, kiss_fft.c:250-286
        int in = 0, stride = 1, rem = o;
        for (int j = 0; j < nf; ++j) {
            const int dgt = rem / fm[j];
            rem -= dgt * fm[j];
            in += dgt * stride;
            stride *= fp[j];
        }
        t.perm[o] = (uint16_t)in;
    }
    t.tw_fwd.resize(2 * nc); t.tw_inv.resize(2 * nc); t.stw_fwd.resize(2 * nc); t.stw_inv.resize(2 * nc);
    for (int i = 0; i < nc; ++i) {  // kiss_fft.c:341-347
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / nc;
        t.tw_fwd[2 * i] = (float)cos(phase); t.tw_fwd[2 * i + 1] = (float)sin(phase);
        phase *= -1;
        t.tw_inv[2 * i] = (float)cos(phase); t.tw_inv[2 * i + 1] = (float)sin(phase);
    }
    for (int i = 0; i < nc; ++i) {  // kiss_fftr.c:57-63
        double phase = -3.14159265358979323846264338327 * ((double)i / nc + .5);
        t.stw_fwd[2 * i] = (float)cos(phase); t.stw_fwd[2 * i + 1] = (float)sin(phase);
        phase *= -1;
        t.stw_inv[2 * i] = (float)cos(phase); t.stw_inv[2 * i + 1] = (float)sin(phase);
    }
    t.omega.resize(nc);
    const size_t uhop = (size_t)hop, uN = (size_t)N;
    for (int i = 0; i < nc; ++i) t.omega[i] = (float)((2 * M_PI * uhop * i) / (uN));
    return t;
}

// glibc 2.39 random_r TYPE_3 (degree 31, separation 3), srandom(1), output word >> 1.
GlibcRand::GlibcRand() {
    int32_t word = 1;
    r_[0] = 1;
    for (int i = 1; i < 31; ++i) {
        const long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        r_[i] = word;
    }
    f_ = 3;
    b_ = 0;
    for (int i = 0; i < 310; ++i) (void)next();
}

int32_t GlibcRand::next() {
    const uint32_t v = (uint32_t)r_[f_] + (uint32_t)r_[b_];
    r_[f_] = (int32_t)v;
    if (++f_ >= 31) f_ = 0;
    if (++b_ >= 31) b_ = 0;
    return (int32_t)(v >> 1);
}

void glibc_rand_fresh(int32_t *dst, size_t n) {
    GlibcRand g;
    for (size_t i = 0; i < n; ++i) dst[i] = g.next();
}

// rosenberg.cc:19-53
Carrier::Pulse Carrier::make(float sr, float freq, float alpha, float beta) {
    Pulse g;
    g.period = (int)round(1.f / freq * sr);
    g.phase = 0;
    g.n1 = (int)round(alpha * g.period);
    g.inv_n1 = 1.f / (float)g.n1;
    g.n2 = (int)round(beta * g.period);
    g.inv_2n2 = 0.5 / (float)g.n2;
    return g;
}

float Carrier::Pulse::next() {
    float res;
    if (phase <= n1) res = 0.5 * (1 - cosf(M_PI * phase * inv_n1));
    else if (phase - n1 <= n2) res = cosf(M_PI * (phase - n1) * inv_2n2);
    else res = 0;
    if (++phase > period) phase = 0;
    return res;
}

Carrier::Carrier(int sample_rate, bool chord) : chord_(chord) {  // impl.cc:312-320
    g_[0] = make((float)sample_rate, 440, 0.01, 0.06);
    g_[1] = make((float)sample_rate, 523.251, 0.01, 0.06);
    g_[2] = make((float)sample_rate, 659.255, 0.01, 0.06);
}

void Carrier::generate(float *dst, size_t n) {
    if (!chord_) {
        for (size_t i = 0; i < n; ++i) dst[i] = g_[0].next() * 0.3;
    } else {
        for (size_t i = 0; i < n; ++i) {
            float res = 0;
            res += g_[0].next() / 3;
            res += g_[1].next() / 3;
            res += g_[2].next() / 3;
            dst[i] = res * 0.3;
        }
    }
}

void carrier_signal(int sample_rate, bool chord, float *dst, size_t n) {
    Carrier c(sample_rate, chord);
    c.generate(dst, n);
}

// ---------------------------------------------------------------------------------------
// scheduler
// ---------------------------------------------------------------------------------------
Scheduler::Scheduler(const Derived &d, bool track_norm) : d_(d), track_norm_(track_norm) {
    if (track_norm_) {
        Tables t = make_tables(d.N, d.hop);
        window_ = t.window;
        acc_scale_ = t.acc_scale;
        winacc_.assign(2 * (size_t)d.N, 0.f);
        winacc_[0] = 1.f;  // channelinfo.cc:108
    }
}

int Scheduler::feed(long n) {
    if (!d_.valid_mode) return 0;
    const size_t before = recs_.size();
    long done = 0;
    bool allread = false;
    while (!allread) {  // Impl::processNormal / processVocoder / processConstant, impl.cc:340-423
        long w = n - done;
        const long space = 2L * d_.N - fill_;
        if (w > space) w = space;
        if (w < 0) w = 0;
        fill_ += w;
        done += w;
        in_total_ += w;
        allread = !(done < n);
        one_slice();
    }
    return (int)(recs_.size() - before);
}

void Scheduler::one_slice() {
    const int N = d_.N, hop = d_.hop;
    if (fill_ < N) return;
    fill_ -= hop;
    SliceRec r;
    std::memset(&r, 0, sizeof(r));
    // increments: processOneSlice, phasevocoderprocess.cc:265-277
    long phase_inc, shift_inc;
    if (d_.vocoder || d_.constant_mode || d_.robotic || d_.whisper) {
        phase_inc = shift_inc = hop;
    } else if (d_.int_ratio) {
        phase_inc = shift_inc = (long)(size_t)(hop * d_.hs);
    } else {  // calculateThisIncrement, phasevocoderprocess.cc:379-410
        const size_t increment = (size_t)hop, samplerate = (size_t)d_.cfg.sample_rate;
        const float ratio = d_.hs;
        recovery_ = divergence_ / ((samplerate / 10.0) / increment);
        int incr = (int)lrint(increment * ratio - recovery_);
        if (incr < lrint((increment * ratio) / 2)) incr = (int)lrint((increment * ratio) / 2);
        else if (incr > lrint(increment * ratio * 2)) incr = (int)lrint(increment * ratio * 2);
        const float divdiff = (increment * ratio) - incr;
        const float prev_div = divergence_;
        divergence_ -= divdiff;
        if ((prev_div < 0 && divergence_ > 0) || (prev_div > 0 && divergence_ < 0))
            recovery_ = divergence_ / ((samplerate / 10.0) / increment);
        shift_inc = incr;
        phase_inc = prev_inc_ == 0 ? shift_inc : prev_inc_;
        prev_inc_ = shift_inc;
    }
    r.phase_inc = (int32_t)phase_inc;
    r.shift_inc = (int32_t)shift_inc;
    r.ola_off = ola_total_;
    r.res_off = res_total_;
    r.out_off = out_total_;
    r.rs_last = rs_last_;
    r.rs_frac = rs_frac_;
    // frames still overlapping the write head
    const long k = recs_base_ + (long)recs_.size();
    if (frame_off_.empty()) frame_off_first_ = k;
    frame_off_.push_back(ola_total_);
    {
        size_t drop = 0;
        while (drop < frame_off_.size() && frame_off_[drop] + N <= ola_total_) ++drop;
        if (drop) { frame_off_.erase(frame_off_.begin(), frame_off_.begin() + (long)drop); frame_off_first_ += (long)drop; }
    }
    r.jlo = (int32_t)frame_off_first_;
    // synthesis always accumulates the window (phasevocoderprocess.cc:1073)
    if (track_norm_)
        for (int i = 0; i < N; ++i) winacc_[i] += window_[i] * acc_scale_;
    // ring space check: processSliceForChannel :337-364 (vocoder/constant: space < hop)
    const long space = d_.outbuf_cap - out_fill_;
    const bool simple_path = d_.vocoder || d_.constant_mode;
    const long required = simple_path ? hop : (long)(int)(shift_inc / d_.pitch_scale) + 1;
    if (space < required) {
        r.flags = 1;
        ++dropped_;
        recs_.push_back(r);
        return;
    }
    // writeSlice :1140-1194
    if (track_norm_) {
        norm_.insert(norm_.end(), winacc_.begin(), winacc_.begin() + shift_inc);
        std::memmove(winacc_.data(), winacc_.data() + shift_inc, sizeof(float) * (size_t)(N - shift_inc));
        std::memset(winacc_.data() + (N - shift_inc), 0, sizeof(float) * (size_t)shift_inc);
    }
    long n_res;
    const bool resample = d_.rs.active && !d_.vocoder;  // writeSliceCarrier never resamples (:1196-1231)
    if (resample) {
        const ResamplerSpec &rs = d_.rs;
        if (rs_initial_) {  // speex_resampler_skip_zeros after the first setratio (resampler.cc:766-769)
            rs_last_ = (int)rs.filt_len / 2;
            rs_initial_ = false;
            r.rs_last = rs_last_;
        }
        int in_len = (int)shift_inc;
        const int out_len = (int)lrintf(ceilf((int)shift_inc * rs.ratio));
        int last = rs_last_, out = 0;
        uint32_t frac = rs_frac_;
        while (!(last >= in_len || out >= out_len)) {  // resample.c:462-560
            ++out;
            last += rs.int_adv;
            frac += (uint32_t)rs.frac_adv;
            if (frac >= rs.den) { frac -= rs.den; ++last; }
        }
        if (last < in_len) in_len = last;  // resample.c:1040-1046
        last -= in_len;
        rs_last_ = last;
        rs_frac_ = frac;
        r.consumed = in_len;
        n_res = out;
    } else {
        r.consumed = (int32_t)shift_inc;
        n_res = shift_inc;
    }
    r.n_res = (int32_t)n_res;
    const long n_write = n_res < space ? n_res : space;  // circularqueue::write truncates
    r.n_write = (int32_t)n_write;
    out_fill_ += n_write;
    out_total_ += n_write;
    ola_total_ += shift_inc;
    res_total_ += r.consumed;
    recs_.push_back(r);
}

void Scheduler::trim(long first_slice_kept, long first_ola_kept) {
    if (first_slice_kept > recs_base_) {
        long drop = first_slice_kept - recs_base_;
        if (drop > (long)recs_.size()) drop = (long)recs_.size();
        recs_.erase(recs_.begin(), recs_.begin() + drop);
        recs_base_ += drop;
    }
    if (track_norm_ && first_ola_kept > norm_base_) {
        long drop = first_ola_kept - norm_base_;
        if (drop > (long)norm_.size()) drop = (long)norm_.size();
        norm_.erase(norm_.begin(), norm_.begin() + drop);
        norm_base_ += drop;
    }
}

StreamPlan plan_stream(Scheduler &s, long n_in, int block) {
    const Derived &d = s.derived();
    StreamPlan p;
    p.n_in = n_in;
    if (block <= 0) block = d.cfg.sample_rate / 100 < 480 ? 480 : d.cfg.sample_rate / 100;  // main.cc:149
    long produced = 0, fed = 0;
    for (long i = 0; i < n_in; i += block) {
        const long m = (n_in - i) < block ? (n_in - i) : block;
        s.feed(m);
        fed += m;
        const long k = s.available();
        s.drain(k);
        produced += k;
    }
    if (d.cfg.mode != kStretch && d.valid_mode) {  // main.cc:492-509
        while (produced < n_in) {
            s.feed(block);
            fed += block;
            long k = s.available();
            s.drain(k);
            if (n_in - produced <= k) k = n_in - produced;
            produced += k;
        }
    }
    p.n_out = produced;
    p.n_slices = s.recs_base() + s.slices();
    p.n_fed = fed;
    return p;
}

void partition_streams(const int64_t *n_in, int n_streams, int n_dev, int *owner) {
    if (n_dev < 1) n_dev = 1;
    bool equal = true;
    for (int s = 1; s < n_streams && equal; ++s) equal = n_in[s] == n_in[0];
    if (equal) {
        const int base = n_streams / n_dev, extra = n_streams % n_dev;
        int s = 0;
        for (int d = 0; d < n_dev; ++d)
            for (int i = 0; i < base + (d < extra ? 1 : 0); ++i) owner[s++] = d;
        return;
    }
    std::vector<int> order(n_streams);
    for (int s = 0; s < n_streams; ++s) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return n_in[a] > n_in[b]; });
    for (int i = 0; i < n_streams; ++i) {
        const int lap = i / n_dev, pos = i % n_dev;
        owner[order[i]] = (lap % 2 == 0) ? pos : n_dev - 1 - pos;
    }
}

}  // namespace pvgpu
