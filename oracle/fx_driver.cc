// TEST INFRASTRUCTURE ONLY (oracle/): drives the UNMODIFIED reference's gain / compressor / limiter objects
// (src/gain/gain.cc, src/dynamics/{compressor,limiter}.cc) over a planar float32 file in blocks of 480 samples, in place,
// the way an SDK user chains them behind the phase vocoder (README.md:78-94).  Checker of the post-chain (pv_post.cu).
//
// usage: fxref_drv sr ch in.f32 out.f32 { gain G | compressor THR RATIO MAKEUP ATT REL | limiter THR MAKEUP ATT REL |
//                                          biquad TYPE CUTOFF Q GAIN | equalizer default | equalizer P0 ... P31 } ...
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "audiomod.h"

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s sr ch in.f32 out.f32 effect args...\n", argv[0]); return 2; }
    const int sr = atoi(argv[1]), ch = atoi(argv[2]);
    FILE *fi = fopen(argv[3], "rb");
    if (!fi) { perror(argv[3]); return 1; }
    fseek(fi, 0, SEEK_END);
    const long n = ftell(fi) / 4 / ch;
    fseek(fi, 0, SEEK_SET);
    std::vector<std::vector<float>> x(ch, std::vector<float>(n));
    for (int c = 0; c < ch; ++c) if (fread(x[c].data(), 4, n, fi) != (size_t)n) return 1;
    fclose(fi);
    std::vector<std::unique_ptr<modbase>> chain;
    for (int i = 5; i < argc;) {
        const std::string k = argv[i];
        auto f = [&](int j) { return (float)atof(argv[i + j]); };
        if (k == "gain" && i + 1 < argc) { chain.emplace_back(new gain(sr, ch, f(1))); i += 2; }
        else if (k == "compressor" && i + 5 < argc) { chain.emplace_back(new compressor(sr, ch, f(1), f(2), f(3), f(4), f(5))); i += 6; }
        else if (k == "limiter" && i + 4 < argc) { chain.emplace_back(new limiter(sr, ch, f(1), f(2), f(3), f(4))); i += 5; }
        else if (k == "biquad" && i + 4 < argc) { chain.emplace_back(new biquadfilter(sr, ch, (biquadfilter::Type)atoi(argv[i + 1]), f(2), f(3), f(4))); i += 5; }
        else if (k == "equalizer" && i + 1 < argc && std::string(argv[i + 1]) == "default") { chain.emplace_back(new equalizer(sr, ch)); i += 2; }
        else if (k == "equalizer" && i + 32 < argc) {
            float pl[32];
            for (int j = 0; j < 32; ++j) pl[j] = f(1 + j);
            chain.emplace_back(new equalizer(sr, ch, pl));
            i += 33;
        }
        else { fprintf(stderr, "bad effect spec at '%s'\n", argv[i]); return 2; }
    }
    const int B = 480;
    std::vector<float *> ptr(ch);
    for (long i = 0; i < n; i += B) {
        const int m = (int)(n - i < B ? n - i : B);
        for (int c = 0; c < ch; ++c) ptr[c] = x[c].data() + i;
        for (auto &fx : chain) fx->processBlock(ptr.data(), m);
    }
    FILE *fo = fopen(argv[4], "wb");
    if (!fo) { perror(argv[4]); return 1; }
    for (int c = 0; c < ch; ++c) fwrite(x[c].data(), 4, n, fo);
    fclose(fo);
    return 0;
}
