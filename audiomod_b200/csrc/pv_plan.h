// Host-side plan for the batched phase vocoder: derived sizes, constant tables and the
// data-INDEPENDENT slice schedule.  Pure C++ (no CUDA) so it is unit-tested on CPU.
//
// Everything here is bookkeeping the reference does inline while it streams
// (file:line relative to the reference tree):
//   sizes            src/phasevocoder/phasevocoderimpl.cc:169-263, phasevocoder.cc:24-60
//   block->slices    phasevocoderimpl.cc:340-369 + src/common/base/circularqueue.h (capacity 2N)
//   increments       phasevocoderprocess.cc:265-277, 379-489
//   window sum       phasevocoderprocess.cc:1057-1073, 1152, 1185-1190 (windowAccumulator)
//   resampler clock  src/common/dsp/resampler.cc:740-817, src/common/speex/resample.c:462-560,986-1059
//   output ring      phasevocoderprocess.cc:337-364 (slice dropped when the ring is full)
// None of it depends on the audio, so one schedule serves every stream of a batch and the
// device keeps only data-dependent state.
//
// Coordinates used by the device kernels (all per channel row, all data independent):
//   input sample  n        : frame k analyses input [k*hop, k*hop+N)
//   OLA position  t        : frame k is overlap-added at [ola_off_k, ola_off_k+N); the first
//                            shift_k positions are final after slice k and are normalised by
//                            norm[t] (the window accumulator value the reference divides by)
//   resampler in  u        : slice k appends its first consumed_k normalised samples at res_off_k
//   output        q        : slice k writes n_write_k samples at out_off_k
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace pvgpu {

enum Mode { kConstant = -1, kShift = 0, kGender = 1, kFormant = 2, kVocRosen = 3, kVocChord = 4, kStretch = 5, kRobotic = 6, kWhisper = 7,
            // extensions (not reference modes): gender / formant with the cepstral envelope routine the reference keeps commented out
            kGenderCepstral = 8, kFormantCepstral = 9 };

struct Config {
    int sample_rate = 44100, channels = 1;
    float time_ratio = 1.f, pitch_semitones = 0.f;
    int mode = kShift, coremode = 1, fftsize = 2048, hopsize = 0;
};

// Per-slice record consumed by the kernels (64 bytes).
struct SliceRec {
    int32_t phase_inc;    // phaseIncrement
    int32_t shift_inc;    // shiftIncrement (accumulator advance)
    int32_t n_res;        // resampler outputs this slice (== shift_inc when no resampling)
    int32_t n_write;      // of which stored (ring-space truncation, circularqueue::write)
    int32_t consumed;     // normalised samples the resampler consumed (resample.c:1040-1046)
    int32_t flags;        // bit0: slice dropped (synthesised into the accumulators, nothing written)
    int32_t rs_last;      // resampler last_sample at slice start
    uint32_t rs_frac;     // resampler samp_frac_num at slice start
    int64_t ola_off;      // OLA position of this frame
    int64_t res_off;      // resampler-input position of this slice's first sample
    int64_t out_off;      // output position (samples per channel)
    int32_t jlo;          // first frame whose OLA span reaches ola_off
    int32_t pad;
};
static_assert(sizeof(SliceRec) == 64, "SliceRec layout");

struct ResamplerSpec {
    bool active = false;
    uint32_t num = 1, den = 1, filt_len = 64, oversample = 8;
    int int_adv = 1, frac_adv = 0;
    bool direct = false;
    float cutoff = 0.94f, ratio = 1.f;
    std::vector<float> table;  // interpolated: filt_len*oversample+8 ; direct: filt_len*den
};

struct Derived {
    Config cfg;
    int N = 2048, H = 1025, hop = 256;
    float pitch_scale = 1.f, hs = 1.f;
    long outbuf_cap = 0;
    bool robotic = false, whisper = false, formant = false, gender = false, vocoder = false, constant_mode = false;
    bool valid_mode = true;
    bool int_ratio = false;
    float freq_comp = 0.f;   // 0: no frequency-axis warp
    bool cepstral = false;   // modes 8 / 9: formantShiftSlice (phasevocoderprocess.cc:925-999) instead of freqCompSlice
    float env_comp = 1.f;    //   its envelope warp factor (:824-840: 1 formant, 0.85 male->female, 1.17 female->male)
    float fixed_gain = 1.f;
    ResamplerSpec rs;
};

Derived derive(const Config &cfg);

// Constant tables (host libm, same expressions as the reference so the floats are identical).
struct Tables {
    std::vector<float> window;      // N, periodic Hann (windowfunc.h:101-169)
    float window_area = 0.f;        // GetArea()
    float acc_scale = 0.f;          // area * 1.5 (phasevocoderprocess.cc:1073)
    std::vector<float> tw_fwd;      // 2*(N/2): re,im  (kiss_fft.c:341-347)
    std::vector<float> tw_inv;
    std::vector<float> stw_fwd;     // 2*(N/2): re,im  (kiss_fftr.c:57-63)
    std::vector<float> stw_inv;
    std::vector<uint16_t> perm;     // N/2: input index landing at slot o (kiss_fft.c:250-286)
    std::vector<int> radix, span;   // execution order (innermost recursion level first)
    std::vector<float> omega;       // N/2: (float)((2*M_PI*hop*i)/N)  (phasevocoderprocess.cc:625)
};
Tables make_tables(int N, int hop);

// glibc rand() of a fresh process (whisperSlice, phasevocoderprocess.cc:820).
class GlibcRand {
public:
    GlibcRand();
    int32_t next();
private:
    int32_t r_[31];
    int f_, b_;
};
void glibc_rand_fresh(int32_t *dst, size_t n);
// carrier pulse trains (rosenberg.cc:19-53, rosenbergchord.cc:38-43, *0.3 at process.cc:100,105)
class Carrier {
public:
    Carrier(int sample_rate, bool chord);
    void generate(float *dst, size_t n);   // the next n samples
private:
    struct Pulse { int period, n1, n2, phase; float inv_n1, inv_2n2; float next(); };
    static Pulse make(float sr, float freq, float alpha, float beta);
    bool chord_;
    Pulse g_[3];
};
void carrier_signal(int sample_rate, bool chord, float *dst, size_t n);

// Incremental scheduler: feed block sizes, get slice records.  One per instance/batch shape.
class Scheduler {
public:
    explicit Scheduler(const Derived &d, bool track_norm = true);
    // One processInData/processBlock call of n samples per channel.  Appends the slices that
    // call runs to recs()/norm(); returns how many were appended.
    int feed(long n);
    // Samples available in the output ring (min over channels == same for all channels).
    long available() const { return out_fill_; }
    // getOutData / retrieve: remove k samples from the ring.
    void drain(long k) { out_fill_ -= (k < out_fill_ ? k : out_fill_); }
    const std::vector<SliceRec> &recs() const { return recs_; }
    const std::vector<float> &norm() const { return norm_; }   // indexed by OLA position
    long slices() const { return (long)recs_.size(); }
    long total_out() const { return out_total_; }     // samples written to the ring so far
    long ola_total() const { return ola_total_; }
    long res_total() const { return res_total_; }
    long in_total() const { return in_total_; }
    long dropped() const { return dropped_; }
    const Derived &derived() const { return d_; }
    // forget records/norm entries older than the given slice / OLA position (streaming use)
    void trim(long first_slice_kept, long first_ola_kept);
    long recs_base() const { return recs_base_; }
    long norm_base() const { return norm_base_; }

private:
    void one_slice();
    Derived d_;
    bool track_norm_;
    std::vector<float> window_;
    float acc_scale_ = 0.f;
    long fill_ = 0, in_total_ = 0;
    long out_fill_ = 0, out_total_ = 0, ola_total_ = 0, res_total_ = 0, dropped_ = 0;
    float recovery_ = 0.f, divergence_ = 0.f;
    long prev_inc_ = 0;
    bool rs_initial_ = true;
    int rs_last_ = 0;
    uint32_t rs_frac_ = 0;
    std::vector<float> winacc_;
    std::vector<SliceRec> recs_;
    std::vector<float> norm_;
    std::vector<long> frame_off_;   // ola_off of every frame still overlapping the write head
    long frame_off_first_ = 0;      // frame index of frame_off_[0]
    long recs_base_ = 0, norm_base_ = 0;
};

// Whole-stream schedule with the reference CLI's block protocol (main/main.cc:149,471-509):
// blocks of max(480, sr/100), pitch modes flushed with zero blocks until out >= in and
// truncated to the input length; time_stretch not flushed.
struct StreamPlan {
    long n_in = 0, n_out = 0, n_slices = 0, n_fed = 0;
};
StreamPlan plan_stream(Scheduler &s, long n_in, int block /*0 = CLI default*/);

// Streams -> devices: a stream (all of its channels) shares nothing with any other stream.  Equal-length
// batches take a contiguous, balanced block partition; ragged batches are sorted by length (longest first, stable) and
// dealt boustrophedon so that every device gets a similar amount of audio.  owner[s] = index of the device of stream s.
void partition_streams(const int64_t *n_in, int n_streams, int n_dev, int *owner);

}  // namespace pvgpu
