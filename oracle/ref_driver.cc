// TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference library at its
// float32 boundary, one OS process per stream (the reference keeps process-global
// state: phasevocoderprocess.cc:380-384,602,716; phasevocoderimpl.h:231-238).
//
// It follows the block protocol of the reference CLI (main/main.cc:149,471-509):
//   block = max(480, sr/100); processInData(block) -> getOutData(getOutSamples());
//   pitch modes: keep feeding zero blocks until out >= in length, truncate to length;
//   time_stretch (mode 5): no tail flush.
// "rt" protocol = the in-place SDK loop (main.cc:562-571, README.md:78-94):
//   processBlock(block); keep the block only if outputReady().
//
// usage: pvref_drv sr ch ratio semitones mode coremode fftsize in.f32 out.f32 [block] [offline|rt]
// in.f32 / out.f32: planar float32, channel 0 then channel 1, native endian.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include "audiomod.h"

int main(int argc, char **argv) {
    if (argc < 10) {
        fprintf(stderr, "usage: %s sr ch ratio semitones mode coremode fftsize in.f32 out.f32 [block] [offline|rt]\n", argv[0]);
        return 2;
    }
    const int sr = atoi(argv[1]), ch = atoi(argv[2]);
    const float ratio = (float)atof(argv[3]), semis = (float)atof(argv[4]);
    const int mode = atoi(argv[5]), coremode = atoi(argv[6]), fftsize = atoi(argv[7]);
    const char *inpath = argv[8], *outpath = argv[9];
    int block = sr / 100 < 480 ? 480 : sr / 100;
    if (argc > 10 && atoi(argv[10]) > 0) block = atoi(argv[10]);
    const bool rt = argc > 11 && std::string(argv[11]) == "rt";

    FILE *fi = fopen(inpath, "rb");
    if (!fi) { perror(inpath); return 1; }
    fseek(fi, 0, SEEK_END);
    const long bytes = ftell(fi);
    fseek(fi, 0, SEEK_SET);
    const long n = bytes / 4 / ch;
    std::vector<std::vector<float>> in(ch, std::vector<float>(n));
    for (int c = 0; c < ch; ++c)
        if (fread(in[c].data(), 4, n, fi) != (size_t)n) { fprintf(stderr, "short read\n"); return 1; }
    fclose(fi);

    // the library chats on stdout/stderr for every slice; keep the harness quiet
    if (!getenv("PVREF_VERBOSE")) {
        if (!freopen("/dev/null", "w", stdout)) return 1;
        if (!freopen("/dev/null", "w", stderr)) return 1;
    }

    audiomod::phasevocoder pv(sr, ch, ratio, semis, mode, coremode, fftsize);
    modbase *rtbase = &pv;
    modbase_offline *off = &pv;

    std::vector<std::vector<float>> out(ch);
    std::vector<std::vector<float>> buf(ch, std::vector<float>(block));
    // the CLI's out buffer is 4*block (main.cc:161); size generously, we only want the data
    std::vector<std::vector<float>> obuf(ch, std::vector<float>((size_t)block * 64 + 65536));
    std::vector<float *> bp(ch), op(ch);
    for (int c = 0; c < ch; ++c) { bp[c] = buf[c].data(); op[c] = obuf[c].data(); }

    if (rt) {
        for (long i = 0; i < n; i += block) {
            const int m = (int)((n - i) < block ? (n - i) : block);
            for (int c = 0; c < ch; ++c) memcpy(bp[c], in[c].data() + i, 4 * (size_t)m);
            rtbase->processBlock(bp.data(), m);
            if (rtbase->outputReady())
                for (int c = 0; c < ch; ++c) out[c].insert(out[c].end(), bp[c], bp[c] + m);
        }
    } else {
        long produced = 0;
        for (long i = 0; i < n; i += block) {
            const int m = (int)((n - i) < block ? (n - i) : block);
            for (int c = 0; c < ch; ++c) memcpy(bp[c], in[c].data() + i, 4 * (size_t)m);
            off->processInData(bp.data(), m);
            const int k = off->getOutSamples();
            off->getOutData(op.data(), k);
            for (int c = 0; c < ch; ++c) out[c].insert(out[c].end(), op[c], op[c] + k);
            produced += k;
        }
        if (mode != NORMAL_STRETCH) {
            for (int c = 0; c < ch; ++c) memset(bp[c], 0, 4 * (size_t)block);
            while (produced < n) {
                off->processInData(bp.data(), block);
                int k = off->getOutSamples();
                off->getOutData(op.data(), k);
                if (n - produced <= k) k = (int)(n - produced);
                for (int c = 0; c < ch; ++c) out[c].insert(out[c].end(), op[c], op[c] + k);
                produced += k;
            }
        }
    }

    FILE *fo = fopen(outpath, "wb");
    if (!fo) return 1;
    for (int c = 0; c < ch; ++c) fwrite(out[c].data(), 4, out[c].size(), fo);
    fclose(fo);
    return 0;
}
