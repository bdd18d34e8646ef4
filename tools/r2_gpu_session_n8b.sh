# second 8-GPU session of round 2 (final code): bench + reference arm + the one-process multi-device calls
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err
python bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/r2_ref_n8b.json 2> gpurun_out/r2_ref_n8b.err
python - > gpurun_out/r2_one_process_8dev_b.txt 2>&1 <<'PY'
import time, sys
sys.path.insert(0, '.')
import __graft_entry__ as e
pkg = e.load_package(); F = pkg.F
ac = pkg.AcGpu(); assert ac.ac_init(pkg.AC_CUDA) == 1
nd = ac.lib.acgpu_device_count()
w, h = 1920, 1080
sfb, dfb = F.frame_bytes(F.IMG_YUV420P, w, h), F.frame_bytes(F.IMG_RGB24, w, h)
for ndev in (1, 2, 4, 8):
    if ndev > nd: break
    eb = 96 * ndev
    hs, hd = ac.pinned(eb * sfb), ac.pinned(eb * dfb)
    hs.array[:] = 0x80
    for _ in range(2): ac._ok(ac.lib.acgpu_imgconvert_frames_host_multi(hs.ptr, F.IMG_YUV420P, hd.ptr, F.IMG_RGB24, w, h, eb, ndev))
    t0 = time.perf_counter()
    for _ in range(5): ac._ok(ac.lib.acgpu_imgconvert_frames_host_multi(hs.ptr, F.IMG_YUV420P, hd.ptr, F.IMG_RGB24, w, h, eb, ndev))
    print("acgpu_imgconvert_frames_host_multi devices", ndev, "frames/s %.0f" % (5 * eb / (time.perf_counter() - t0)), flush=True)
    hs.free(); hd.free()
# config 4 as "8 streams, 1 per GPU" from ONE process: a UHD round-trip chain over all devices
w, h = 3840, 2160
sfb, dfb = F.frame_bytes(F.IMG_YUV420P, w, h), F.frame_bytes(F.IMG_YUV422P, w, h)
ops = pkg.chain_ops([(pkg.CHAIN_CONVERT, F.IMG_RGB24), (pkg.CHAIN_CONVERT, F.IMG_YUV422P)])
for ndev in (1, 8):
    if ndev > nd: break
    eb = 24 * ndev
    hs, hd = ac.pinned(eb * sfb), ac.pinned(eb * dfb)
    hs.array[:] = 0x80
    for _ in range(2): ac._ok(ac.lib.acgpu_chain_frames_host_multi(hs.ptr, F.IMG_YUV420P, w, h, hd.ptr, ops, 2, eb, ndev))
    t0 = time.perf_counter()
    for _ in range(5): ac._ok(ac.lib.acgpu_chain_frames_host_multi(hs.ptr, F.IMG_YUV420P, w, h, hd.ptr, ops, 2, eb, ndev))
    print("acgpu_chain_frames_host_multi (UHD 420P->RGB24->422P) devices", ndev, "frames/s %.0f" % (5 * eb / (time.perf_counter() - t0)), flush=True)
    hs.free(); hd.free()
PY
