"""Dev: when do the persistent warps of a traversal launch run dry / thin out / exit?

Build the instrumented library (`make -C tray_racing_b200/csrc exitlog`), then on a GPU
  TRAY_CUDA_LIB=$PWD/tray_racing_b200/libtray_cuda_exitlog.so TRAY_EXIT_LOG_FILE=gpurun_out/exitlog \
      python scripts/render_frames.py --scene hairball --frames 3
and read the dumps here:  python scripts/exit_log.py gpurun_out/exitlog.*.bin
Each dump holds, per warp, the %globaltimer (ns) at which the work cursor was exhausted for it, at which it exited, and at
which it was first down to <= 8 / 4 / 2 / 1 rays after the cursor ran dry."""
import sys

import numpy as np

for f in sorted(sys.argv[1:], key=lambda x: int(x.split(".")[-2])):
    a = np.fromfile(f, dtype=np.uint64).reshape(-1, 6).astype(np.int64)
    dry, ex = a[:, 0], a[:, 1]
    ok = ex > 0
    t_end = ex[ok].max()
    d = dry[ok & (dry > 0)]
    q = (t_end - np.percentile(ex[ok], [10, 50, 90, 99])) / 1e3
    print(f"{f}: {int(ok.sum())} warps | cursor dry {(t_end - np.median(d)) / 1e3:.1f} us before the launch ends (first {(t_end - d.min()) / 1e3:.1f}, last {(t_end - d.max()) / 1e3:.1f})"
          f" | warps exit p10 {q[0]:.1f} p50 {q[1]:.1f} p90 {q[2]:.1f} p99 {q[3]:.1f} us before the end")
    # the warps that end the launch: how long did each spend with few rays left?
    last = np.argsort(-ex)[: max(1, int(ok.sum()) // 100)]
    for name, col in (("<= 8", 2), ("<= 4", 3), ("<= 2", 4), ("<= 1", 5)):
        t = a[last, col]
        v = t > 0
        if v.any():
            print(f"    last 1 % of warps: {name} rays for the final {np.median((ex[last] - t)[v]) / 1e3:6.1f} us (median; {int(v.sum())} of {len(last)} warps got there)")
    # all warps: lane-idle time after the cursor ran dry, as a share of warp-time in the launch
    span = (t_end - dry[ok & (dry > 0)].min()) / 1e3
    for name, col in (("<= 4", 3), ("<= 1", 5)):
        t = a[:, col]; v = ok & (t > 0)
        print(f"    all warps: time spent with {name} rays, summed: {((ex - t)[v]).sum() / 1e3 / max(1, ok.sum()):.1f} us per warp (drain phase lasts {span:.0f} us)")
