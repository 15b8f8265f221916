// C++ host-side mirror of the reference's phase-vocoder class over the C ABI (include/pvgpu.h).
//
// Same names, constructor arguments, virtual interface and error behaviour as the reference's
// include/dafx/modbase.h:25-121 and include/dafx/phasevocoder.h:22-115, so code written against the
// reference (main/main.cc:165-170,196-287,471-509; README SDK loop) compiles and links unchanged:
//     modbase_offline *fx = new audiomod::phasevocoder(sr, ch, 1, 7, NORMAL_SHIFT, PHASE_LOCKED, 2048);
//     fx->processInData(buf, n);  fx->getOutData(out, fx->getOutSamples());
// The object layout (two vptrs, then the private block starting with the engine pointer) is kept identical to the
// reference's so that objects allocated by callers compiled against the reference header have the right size; when
// the reference headers are on the include path the shim can be built against them directly with
// -DPVGPU_DROPIN_USE_REFERENCE_HEADERS.
#pragma once

#ifdef PVGPU_DROPIN_USE_REFERENCE_HEADERS
#include "phasevocoder.h"
#else
#include <map>
#include <string>

// real-time, in-place interface
class modbase {
public:
    modbase() : sample_rate_(48000), num_channels_(1) {}
    virtual ~modbase() {}
    virtual void processBlock(float *const *bufferData, int num_samples) = 0;
    virtual void setParams(std::map<std::string, float> params) = 0;
    virtual void getParams(std::map<std::string, float> &params) = 0;
    virtual bool outputReady() { return true; }

protected:
    int sample_rate_;
    int num_channels_;
};

// offline interface: push input, pull whatever is ready
class modbase_offline {
public:
    modbase_offline() : sample_rate_(48000), num_channels_(1), num_res_(0) {}
    virtual ~modbase_offline() {}
    virtual void processInData(float *const *inData, int num_in_samples) = 0;
    virtual void getOutData(float *const *outData, int num_out_samples) = 0;
    virtual void setParams(std::map<std::string, float> params) = 0;
    virtual void getParams(std::map<std::string, float> &params) = 0;
    virtual int getOutSamples() const { return num_res_; }
    virtual bool outputReady() { return true; }

protected:
    int sample_rate_;
    int num_channels_;
    int num_res_;
};

#define CONSTANT -1
#define NORMAL_SHIFT 0
#define GENDER_CHANGE 1
#define FORMANT_PRESERVE 2
#define VOCODER_ROSENBERG 3
#define VOCODER_CHORD 4
#define NORMAL_STRETCH 5
#define ROBOTIC 6
#define WHISPER 7
#define NORMAL_PV 0
#define PHASE_LOCKED 1
#define INT_RATIO 2

namespace audiomod {
class phasevocodercore;  // opaque: here it is the pvgpu_stream handle

class phasevocoder : public modbase, public modbase_offline {
public:
    phasevocoder(int sampleRate, int numChannels, float timeratio, float pitchshift, int mode = NORMAL_SHIFT,
                 int coremode = PHASE_LOCKED, int fftsize = 2048, int hopsize = 0);
    ~phasevocoder();
    void setParams(std::map<std::string, float>) {}
    void getParams(std::map<std::string, float> &) {}
    void processBlock(float *const *bufferData, int num_samples);
    void processInData(float *const *inData, int num_in_samples);
    void getOutData(float *const *outData, int num_out_samples);
    bool outputReady() { return outready_; }

private:
    phasevocoder(const phasevocoder &) = delete;
    void operator=(const phasevocoder &) = delete;
    void init();
    int processBlockNormal(float *const *bufferData, int num_samples);
    int processBlockConstant(float *const *bufferData, int num_samples);
    int processBlockVocoder(float *const *bufferData, int num_samples, int carrierType);

    phasevocodercore *ts;
    int options_;
    int m_mode;
    int m_log;
    float timeratio_;
    float pitchscale_;
    int defaultfftsize_;
    int defaulthopsize_;
    int defaultcoremode_;
    int sample_rate_;
    int num_channels_;
    bool outready_;
};
}  // namespace audiomod
#endif
