"""Host-side probe for the reference arm: how many cores are really available, and how the reference's fused C2 loop
scales with threads by the wall clock (used to pick the thread count of bench.py --impl reference)."""
import os, sys, time
sys.path.insert(0, '.')
import bench
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
try: print(open("/proc/loadavg").read().strip())
except Exception: pass
for th in (1, 8, 16, 32, 64, 128):
    if th > (os.cpu_count() or 1): break
    c = bench.CpuC2(bench.N_ELEMS, th)
    c.run(1)
    t0 = time.time(); c.run(8); wall = time.time() - t0
    print(f"threads {th:4d}: {8 * c.per * c.threads * bench.C2_NODES / wall / 1e9:8.1f} G array-ops/s by wall clock ({wall / 8 * 1e3:.1f} ms per pass)", flush=True)
