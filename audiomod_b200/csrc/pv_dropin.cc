// audiomod::phasevocoder implemented over the C ABI: see pv_dropin.hpp.  Replaces
// src/phasevocoder/phasevocoder.cc of the reference (facade) -- the engine behind `ts` is the GPU stream.
#include "pv_dropin.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../../include/pvgpu.h"

namespace audiomod {

static inline pvgpu_stream *handle(phasevocodercore *p) { return reinterpret_cast<pvgpu_stream *>(p); }

// PVGPU_DEVICE selects the CUDA device for instances created through the C++ class (default 0).
static int env_device() {
    const char *e = std::getenv("PVGPU_DEVICE");
    return e ? std::atoi(e) : 0;
}

phasevocoder::phasevocoder(int sampleRate, int numChannels, float timeratio, float pitchshift, int mode, int coremode, int fftsize,
                           int hopsize) {
    // phasevocoder.cc:24-60
    timeratio_ = timeratio;
    pitchscale_ = pitchshift != 0 ? std::pow(2.0, pitchshift / 12) : 1.0;
    options_ = 0;
    sample_rate_ = sampleRate;
    num_channels_ = numChannels;
    defaultfftsize_ = fftsize;
    defaulthopsize_ = hopsize;
    defaultcoremode_ = coremode;
    m_mode = mode;
    m_log = 0;
    outready_ = false;
    ts = nullptr;
    modbase_offline::num_res_ = 0;
    pvgpu_config cfg;
    cfg.sample_rate = sampleRate; cfg.channels = numChannels; cfg.time_ratio = timeratio; cfg.pitch_semitones = pitchshift;
    cfg.mode = mode; cfg.coremode = coremode; cfg.fftsize = fftsize; cfg.hopsize = hopsize; cfg.device = env_device();
    pvgpu_stream *s = nullptr;
    if (pvgpu_create(&cfg, &s) != PVGPU_OK) {
        // the reference aborts on construction failures (FFT.cc:3168, memallocators.h:91); here: report, stay inert
        std::fprintf(stderr, "audiomod::phasevocoder (GPU): %s\n", pvgpu_last_error());
        s = nullptr;
    }
    ts = reinterpret_cast<phasevocodercore *>(s);
}

phasevocoder::~phasevocoder() {
    if (ts != nullptr) {
        pvgpu_destroy(handle(ts));
        ts = nullptr;
    }
}

void phasevocoder::init() {}

void phasevocoder::processInData(float *const *inData, int num_in_samples) {  // phasevocoder.cc:87-108
    if (!ts || pvgpu_process(handle(ts), inData, num_in_samples) != PVGPU_OK) {
        if (ts) std::fprintf(stderr, "audiomod::phasevocoder (GPU): %s\n", pvgpu_last_error());
        num_res_ = 0;
        return;
    }
    num_res_ = pvgpu_available(handle(ts));
}

void phasevocoder::getOutData(float *const *outData, int num_out_samples) {  // phasevocoder.cc:110-124
    if (ts) pvgpu_retrieve(handle(ts), outData, num_out_samples);
    outready_ = true;
}

void phasevocoder::processBlock(float *const *bufferData, int num_samples) {  // phasevocoder.cc:126-154
    int ready = 0;
    if (!ts || pvgpu_process_block(handle(ts), bufferData, num_samples, &ready) != PVGPU_OK) {
        if (ts) std::fprintf(stderr, "audiomod::phasevocoder (GPU): %s\n", pvgpu_last_error());
        outready_ = false;
        return;
    }
    num_res_ = pvgpu_available(handle(ts));
    outready_ = ready != 0;
}

int phasevocoder::processBlockNormal(float *const *b, int n) { processBlock(b, n); return outready_ ? 0 : -1; }
int phasevocoder::processBlockConstant(float *const *b, int n) { processBlock(b, n); return outready_ ? 0 : -1; }
int phasevocoder::processBlockVocoder(float *const *b, int n, int) { processBlock(b, n); return outready_ ? 0 : -1; }

}  // namespace audiomod
