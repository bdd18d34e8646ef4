# 8-GPU session: the driver's launch line for the bench (all ranks, extras included), the reference arm on the same box,
# and the host-side DMA probe in every variant.
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python bench.py --impl reference --gpus $N --steps 5 --warmup 1 > gpurun_out/r2_ref_n$N.json 2> gpurun_out/r2_ref_n$N.err
nproc > gpurun_out/r2_host_n$N.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2_host_n$N.txt; free -g | head -2 >> gpurun_out/r2_host_n$N.txt
ls /sys/devices/system/node/ | grep node >> gpurun_out/r2_host_n$N.txt
nvidia-smi topo -m >> gpurun_out/r2_host_n$N.txt 2>&1
if [ "$N" = 8 ]; then
  P=gpurun_out/r2_dma_probe_n8.txt; : > $P
  for g in 1 2 4 8; do
    for dir in h2d d2h both; do ./tools/host_dma_probe --gpus $g --dir $dir --seconds 1.5 >> $P 2>&1; done
  done
  for mode in threads procs; do for mem in hostalloc register thp; do
    ./tools/host_dma_probe --gpus 8 --mode $mode --mem $mem --dir both --seconds 1.5 >> $P 2>&1
  done; done
  ./tools/host_dma_probe --gpus 8 --mode threads --mem hostalloc --dir both --numa --seconds 1.5 >> $P 2>&1
  ./tools/host_dma_probe --gpus 8 --mode procs --mem hostalloc --dir both --numa --seconds 1.5 >> $P 2>&1
  ./tools/host_dma_probe --gpus 8 --mode procs --mem hostalloc --dir both --region 64 --chunk 16 --seconds 1.5 >> $P 2>&1
  ./tools/host_dma_probe --gpus 8 --mode procs --mem hostalloc --dir both --region 2048 --chunk 64 --seconds 1.5 >> $P 2>&1
  ./tools/host_dma_probe --gpus 8 --mode procs --mem hugetlb --dir both --seconds 1.5 >> $P 2>&1
  # the multi-device host calls of ONE process
  python - > gpurun_out/r2_one_process_8dev.txt 2>&1 <<'PY'
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
    eb = 64 * ndev
    hs, hd = ac.pinned(eb * sfb), ac.pinned(eb * dfb)
    hs.array[:] = 0x80
    for _ in range(2): ac._ok(ac.lib.acgpu_imgconvert_frames_host_multi(hs.ptr, F.IMG_YUV420P, hd.ptr, F.IMG_RGB24, w, h, eb, ndev))
    t0 = time.perf_counter()
    for _ in range(5): ac._ok(ac.lib.acgpu_imgconvert_frames_host_multi(hs.ptr, F.IMG_YUV420P, hd.ptr, F.IMG_RGB24, w, h, eb, ndev))
    print("acgpu_imgconvert_frames_host_multi devices", ndev, "frames/s %.0f" % (5 * eb / (time.perf_counter() - t0)), flush=True)
    hs.free(); hd.free()
PY
fi
