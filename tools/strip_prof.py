"""Developer tool: per-CTA ring-wait cycles of the dense strip kernel (build with FUVS_STRIP_PROF=1)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from flood_uav_video_segmentation_b200 import kernels
mode = sys.argv[1] if len(sys.argv) > 1 else "dense"
dev = torch.device("cuda", 0)
clip = bench.make_clip(mode, dev, 0)
counts = kernels.new_counts(bench.C, dev)
lib = kernels.load()
need = int(lib.fuvs_dense_scratch_floats(bench.C, bench.H, bench.W, bench.K_DELTA))
bench.run_interval.scratch = torch.empty((need,), dtype=torch.float32, device=dev)
for _ in range(3):
    bench.run_clip(kernels, mode, clip, counts)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (16 * 148 * 4))()
assert lib.fuvs_debug_strip_prof(buf) == 0
a = np.array(buf[:], dtype=np.float64).reshape(16, 148, 4)
for l in range(16):
    t = a[l]
    print(f"{mode} launch%16={l} (step {(l % 4) + 1}): total cycles mean {t[:,0].mean():.0f} max {t[:,0].max():.0f} min {t[:,0].min():.0f} | "
          f"ring wait mean {t[:,1].mean():.0f} max {t[:,1].max():.0f} | first-block wait mean {t[:,2].mean():.0f} | "
          f"wait share {(t[:,1].sum()+t[:,2].sum())/t[:,0].sum():.3f}")
for l in (5, 6, 7):
    t = a[l]
    order = np.argsort(-t[:, 0])[:12]
    print(f"launch {l}: slowest CTAs (idx,total,ringwait,firstwait,blocks):", [(int(i), int(t[i, 0]), int(t[i, 1]), int(t[i, 2]), int(t[i, 3])) for i in order])
    print("   per-CTA totals by index /1000:", [int(v / 1000) for v in t[:, 0]])
