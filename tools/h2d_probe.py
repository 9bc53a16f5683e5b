"""H2D bandwidth of pinned staging buffers by the NUMA node they were allocated from (run on the GPU box).
Usage: python tools/h2d_probe.py   -> one line per node: GB/s of MoB200_Upload of a 1 GiB pinned buffer."""
import ctypes, glob, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matrixone_b200 import capi

lib = capi.load_library()
capi.check(lib.MoB200_Init(0), lib)
nbytes = 1 << 30
dev = ctypes.c_void_p()
capi.check(lib.MoB200_DeviceAlloc(nbytes, ctypes.byref(dev)), lib)
all_cpus = os.sched_getaffinity(0)
nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
for g in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    try:
        cls = open(os.path.join(os.path.dirname(g), "class")).read().strip()
        if cls.startswith("0x0302") or cls.startswith("0x0300"):
            print("gpu", os.path.basename(os.path.dirname(g)), "numa_node", open(g).read().strip())
    except OSError:
        pass

def cpulist(s):
    out = set()
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out

for nd in nodes + [None]:
    cpus = cpulist(open(nd + "/cpulist").read()) & all_cpus if nd else all_cpus
    if not cpus:
        continue
    os.sched_setaffinity(0, cpus)
    host = ctypes.c_void_p()
    capi.check(lib.MoB200_HostAlloc(nbytes, ctypes.byref(host)), lib)
    ctypes.memset(host, 1, nbytes)
    best = 0.0
    for _ in range(4):
        t0 = time.perf_counter()
        capi.check(lib.MoB200_Upload(dev, host, nbytes), lib)
        capi.check(lib.MoB200_Sync(), lib)
        best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
    print("alloc on", os.path.basename(nd) if nd else "any", "cpus", len(cpus), "H2D GB/s %.1f" % best)
    lib.MoB200_HostFree(host)
os.sched_setaffinity(0, all_cpus)
