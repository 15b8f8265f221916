"""Live batch (pvgpu_create_multi): S phasevocoder objects of one configuration advancing in lock-step through one instance.

Every stream of the live batch must produce, call by call, exactly the samples (and counts) its own single-stream instance
produces for the same sequence of calls -- and those single-stream instances are the ones the golden / oracle tests pin to
the reference (tests/test_gpu_parity.py::test_streaming_*).  Checked bit for bit: the live batch runs the same kernels on
rows = S x channels with the same data-independent schedule."""
import numpy as np
import pytest

from cases import CASES, ctor_args, make_input

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A(pvlib):
    import audiomod_b200
    if pvlib.pvgpu_device_count() < 1:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return audiomod_b200


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


LIVE = ["cfg4_shift_p7_mono", "cfg1_shift_p4_stereo", "cfg3_formant_p4", "cfg5_robotic_512", "cfg5_whisper_1024", "cfg5_chord_4096",
        "cfg2_stretch_1p5_4096", "core0_shift_p3_stereo"]


@pytest.mark.parametrize("name", LIVE)
def test_live_batch_equals_single_instances_call_by_call(A, name):
    kw, sr, ch, secs, seed = next((c[1], c[2], c[3], c[4], c[5]) for c in CASES if c[0] == name)
    tr, st, mode, core, fft = ctor_args(kw)
    S = 5
    n = int(sr * min(secs, 0.5))
    xs = [make_input(name, sr, ch, n / sr, seed + 11 * i)[:, :n] for i in range(S)]
    X = np.ascontiguousarray(np.concatenate(xs, axis=0))           # [S * ch, n], row = stream * ch + channel
    live = A.phasevocoder(sr, ch, tr, st, mode, core, fft, streams=S)
    singles = [A.phasevocoder(sr, ch, tr, st, mode, core, fft) for _ in range(S)]
    sizes = [480, 480, 1, 0, 2048, 311, 4099, 480]
    pos, k, total = 0, 0, 0
    while pos < n:
        m = min(sizes[k % len(sizes)], n - pos)
        k += 1
        live.processInData(np.ascontiguousarray(X[:, pos:pos + m]))
        avail = live.getOutSamples()
        take = avail if k % 3 else avail // 2                       # sometimes leave samples in the ring
        Y = live.getOutData(take)
        assert Y.shape[0] == S * ch
        for i, pv in enumerate(singles):
            pv.processInData(np.ascontiguousarray(xs[i][:, pos:pos + m]))
            assert pv.getOutSamples() == avail, f"{name}: call {k}, stream {i}"
            y = pv.getOutData(take)
            assert np.array_equal(_bits(Y[i * ch:(i + 1) * ch]), _bits(y)), f"{name}: call {k}, stream {i} differs from its own instance"
        total += Y.shape[1]
        pos += m
    assert total > 0
    live.close()
    for pv in singles:
        pv.close()


def test_live_batch_process_block_protocol(A):
    """The in-place real-time protocol (processBlock / outputReady, phasevocoder.cc:133-183) on 64 streams at once."""
    sr, ch, S, B = 44100, 1, 64, 480
    n = B * 60
    xs = [make_input("x", sr, ch, n / sr, 7000 + i)[:, :n] for i in range(S)]
    X = np.ascontiguousarray(np.concatenate(xs, axis=0))
    live = A.phasevocoder(sr, ch, 1.0, 7.0, streams=S)
    ref = [A.phasevocoder(sr, ch, 1.0, 7.0) for _ in (0, S // 2, S - 1)]
    ready_blocks = 0
    for pos in range(0, n, B):
        blk = np.ascontiguousarray(X[:, pos:pos + B])
        before = blk.copy()
        live.processBlock(blk)
        for j, i in enumerate((0, S // 2, S - 1)):
            b1 = np.ascontiguousarray(xs[i][:, pos:pos + B])
            ref[j].processBlock(b1)
            assert ref[j].outputReady() == live.outputReady()
            assert np.array_equal(_bits(blk[i:i + 1]), _bits(b1)), f"block at {pos}, stream {i}"
        if live.outputReady():
            ready_blocks += 1
        else:
            assert np.array_equal(blk, before), "buffers must be untouched while output is not ready"
    assert ready_blocks > 40
    live.close()
    for pv in ref:
        pv.close()


def test_live_batch_limits(A, pvlib):
    import ctypes as C
    from audiomod_b200 import _lib
    with pytest.raises(A.PvgpuError) as e:
        A.phasevocoder(44100, 2, 1.0, 4.0, streams=40000)
    assert e.value.code == _lib.EINVAL
    pv = A.phasevocoder(44100, 2, 1.0, 4.0, streams=3)
    assert pvlib.pvgpu_stream_count(pv._h) == 3
    pv.close()


def test_live_batch_large_uses_the_copy_pool(A):
    """300 rows and more start the row-copy pool (page-locked packing with non-temporal stores on several threads, ring FIFO
    filled by the device-to-host copy): a few streams against their own instances, odd call sizes included, and a call larger
    than the pool's 1 MB threshold."""
    sr, ch, S = 44100, 1, 320
    n = 480 * 25 + 4099
    base = make_input("x", sr, ch, n / sr, 8100)[:, :n]
    X = np.ascontiguousarray(np.repeat(base, S, axis=0) * (1.0 + 0.01 * np.arange(S, dtype=np.float32))[:, None])
    live = A.phasevocoder(sr, ch, 1.0, 7.0, streams=S)
    picks = (0, 1, 77, S - 1)
    singles = [A.phasevocoder(sr, ch, 1.0, 7.0) for _ in picks]
    sizes = [480] * 10 + [4099, 1, 0, 480, 311] + [480] * 12
    pos = 0
    for m in sizes:
        m = min(m, n - pos)
        live.processInData(np.ascontiguousarray(X[:, pos:pos + m]))
        avail = live.getOutSamples()
        Y = live.getOutData(avail)
        for j, i in enumerate(picks):
            singles[j].processInData(np.ascontiguousarray(X[i:i + 1, pos:pos + m]))
            assert singles[j].getOutSamples() == avail
            y = singles[j].getOutData(avail)
            assert np.array_equal(_bits(Y[i:i + 1]), _bits(y)), f"stream {i} at {pos}"
        pos += m
    live.close()
    for pv in singles:
        pv.close()


@pytest.mark.parametrize("name", ["cfg4_shift_p7_mono", "cfg1_shift_p4_stereo", "cfg5_whisper_1024", "cfg5_chord_4096", "cfg2_stretch_1p5_4096"])
def test_device_rows_equal_host_rows(A, name):
    """pvgpu_process_device / _retrieve_device: rows in device memory, everything enqueued on one CUDA stream and no call waiting
    for the device -- same counts and bit-identical samples as the host-row calls, for a live batch and odd call sizes."""
    import torch
    kw, sr, ch, secs, seed = next((c[1], c[2], c[3], c[4], c[5]) for c in CASES if c[0] == name)
    tr, st, mode, core, fft = ctor_args(kw)
    S = 6
    n = int(sr * min(secs, 0.5))
    X = np.ascontiguousarray(np.concatenate([make_input(name, sr, ch, n / sr, seed + 5 * i)[:, :n] for i in range(S)], axis=0))
    R = S * ch
    host = A.phasevocoder(sr, ch, tr, st, mode, core, fft, streams=S)
    dev = A.phasevocoder(sr, ch, tr, st, mode, core, fft, streams=S)
    stream = torch.cuda.Stream()
    dX = torch.from_numpy(X).cuda()
    pitch_out = 1 << 16
    dY = torch.zeros((R, pitch_out), dtype=torch.float32, device="cuda")
    sizes = [480, 480, 1, 0, 2048, 311, 4099, 480]
    pos, k, got_host, counts = 0, 0, [], []
    w = 0
    with torch.cuda.stream(stream):
        while pos < n:
            m = min(sizes[k % len(sizes)], n - pos)
            k += 1
            host.processInData(np.ascontiguousarray(X[:, pos:pos + m]))
            avail = host.getOutSamples()
            take = avail if k % 3 else avail // 2
            got_host.append(host.getOutData(take))
            dev.processInDataDevice(dX.data_ptr() + 4 * pos, X.shape[1], m, stream.cuda_stream)
            assert dev.getOutSamples() == avail, f"{name}: call {k}"
            c = dev.getOutDataDevice(dY.data_ptr() + 4 * w, pitch_out, take, stream.cuda_stream)
            assert c == got_host[-1].shape[1]
            w += c
            pos += m
    stream.synchronize()
    want = np.concatenate(got_host, axis=1)
    assert want.shape[1] == w and w > 0
    assert np.array_equal(_bits(dY[:, :w].cpu().numpy()), _bits(want)), f"{name}: device rows differ from host rows"
    with pytest.raises(A.PvgpuError):
        dev.processInData(np.zeros((R, 16), np.float32))       # an instance takes host rows or device rows, not both
    host.close()
    dev.close()


def test_device_rows_process_block_protocol(A):
    import torch
    sr, S, B = 44100, 48, 480
    n = B * 50
    X = np.ascontiguousarray(np.concatenate([make_input("x", sr, 1, n / sr, 9100 + i)[:, :n] for i in range(S)], axis=0))
    host = A.phasevocoder(sr, 1, 1.0, 7.0, streams=S)
    dev = A.phasevocoder(sr, 1, 1.0, 7.0, streams=S)
    stream = torch.cuda.Stream()
    d = torch.from_numpy(X).cuda()
    ready = []
    with torch.cuda.stream(stream):
        for pos in range(0, n, B):
            dev.processBlockDevice(d.data_ptr() + 4 * pos, n, B, stream.cuda_stream)     # in place, block by block
            ready.append(dev.outputReady())
    stream.synchronize()
    got = d.cpu().numpy()
    for i, pos in enumerate(range(0, n, B)):
        blk = np.ascontiguousarray(X[:, pos:pos + B])
        host.processBlock(blk)
        assert host.outputReady() == ready[i]
        assert np.array_equal(_bits(got[:, pos:pos + B]), _bits(blk)), f"block {i}"
    host.close()
    dev.close()
