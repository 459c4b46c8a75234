"""Developer tool: per-role timeline of the row-tile GEMM kernel on CTA 0 (needs the debug build:
   python 4d-or_b200/build.py --debug;  python tools/trace_mlp.py fwd2|fwd1k|da|sa1fwd [rows])."""
import ctypes, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["SG4D_LIBRARY"] = os.path.join(ROOT, "4d-or_b200", "libsg4d_dbg.so")
sys.path.insert(0, ROOT)
import torch
from sg4d import _lib, mlp
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "fwd2"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 528 * 128 * 64
torch.manual_seed(0)
lib = _lib.load()
trace = torch.zeros(8192, dtype=torch.int64, device=dev)
lib.sg4d_debug_set_trace.argtypes = [ctypes.c_void_p]


def run():
    if which == "fwd2":      # SA2 second layer: 128 -> 128, BN+ReLU prologue, pool 64
        y1 = run.y1 = getattr(run, "y1", None) if getattr(run, "y1", None) is not None else torch.randn(rows, 128, device=dev)
        w = torch.randn(128, 128, device=dev) / 11
        s, t, g = torch.randn(128, device=dev), torch.randn(128, device=dev), torch.randn(128, device=dev)
        return lambda: mlp.linear_fwd(y1, 128, mlp.pack_weight(w), 128, scale=s, shift=t, group=64, gamma=g)
    if which == "fwd1k":     # SA1b second layer (old path): 64 -> 128, pool 32
        y1 = torch.randn(rows, 64, device=dev)
        w = torch.randn(128, 64, device=dev) / 8
        s, t, g = torch.randn(64, device=dev), torch.randn(64, device=dev), torch.randn(128, device=dev)
        return lambda: mlp.linear_fwd(y1, 64, mlp.pack_weight(w), 128, scale=s, shift=t, group=32, gamma=g)
    if which == "da":        # SA2b pooled-layer backward: dY2 (128) * W2 -> dz1 (128), PMODE 2 / EMODE 1
        n1 = n2 = 128
        group = 64
        G = rows // group
        y1, y2 = torch.randn(rows, n1, device=dev), torch.randn(rows, n2, device=dev)
        a2, b2 = torch.randn(n2, device=dev), torch.randn(n2, device=dev)
        p1, q1, u1 = torch.randn(n1, device=dev), torch.randn(n1, device=dev), torch.randn(n1, device=dev)
        dsel = torch.randn(G, n2, device=dev)
        garg = torch.randint(0, group, (G, n2), device=dev, dtype=torch.uint8)
        img = mlp.pack_weight(torch.randn(n1, n2, device=dev))
        dz1 = torch.empty(rows, n1, device=dev)
        part = torch.empty(lib.sg4d_mlp_partial_doubles(rows), dtype=torch.float64, device=dev)
        return lambda: _lib.call("sg4d_pool_bwd_da", y1, rows, n2, n1, group, y2.data_ptr(), a2.data_ptr(), b2.data_ptr(), dsel.data_ptr(),
                                 garg.data_ptr(), img.data_ptr(), y1.data_ptr(), p1.data_ptr(), q1.data_ptr(), p1.data_ptr(), u1.data_ptr(),
                                 dz1.data_ptr(), part.data_ptr())
    raise SystemExit("unknown kernel")


fn = run()
for _ in range(2):
    fn()
torch.cuda.synchronize()
lib.sg4d_debug_set_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fn(); e1.record()
torch.cuda.synchronize()
lib.sg4d_debug_set_trace(None)
print(which, rows, "rows:", round(e0.elapsed_time(e1), 3), "ms")
t = trace.cpu().tolist()


def med(xs):
    xs = [x for x in xs if x is not None]
    return int(statistics.median(xs)) if xs else None


def col(base, stride, k, n):
    return [t[base + i * stride + k] if t[base + i * stride + k] else None for i in range(n)]


def diff(a, b):
    return [y - x if (x and y) else None for x, y in zip(a, b)]


lo, hi = 32, 380
P0, P1, Pa, Pb, P2 = (col(0, 5, k, 400) for k in range(5))
print("producer / k-block: wait-empty", med(diff(P0, P1)[lo:hi]), " transform+STS", med(diff(P1, Pa)[lo:hi]), " fence+syncwarp", med(diff(Pa, Pb)[lo:hi]),
      " arrive", med(diff(Pb, P2)[lo:hi]), " issue-next-loads", med(diff(P2[:-1], P0[1:])[lo:hi]), " period", med(diff(P0[:-1], P0[1:])[lo:hi]))
M0, M1, M2 = col(2048, 3, 0, 400), col(2048, 3, 1, 400), col(2048, 3, 2, 400)
print("mma / k-block:      wait-full", med(diff(M0, M1)[lo:hi]), " issue", med(diff(M1, M2)[lo:hi]), " period", med(diff(M0[:-1], M0[1:])[lo:hi]))
T0, T1 = col(4096, 2, 0, 256), col(4096, 2, 1, 256)
print("mma / tile:         wait-acc-empty", med(diff(T0, T1)[8:120]), " period", med(diff(T0[:-1], T0[1:])[8:120]))
E = [col(5120, 5, k, 256) for k in range(5)]
print("epilogue / tile:    wait-acc-full", med(diff(E[0], E[1])[8:120]), " tmem->smem", med(diff(E[1], E[2])[8:120]),
      " store", med(diff(E[2], E[3])[8:120]), " stats/pool", med(diff(E[3], E[4])[8:120]), " period", med(diff(E[0][:-1], E[0][1:])[8:120]))
# handoff latencies: producer arrive -> mma sees full; mma commit -> producer sees empty (2 stages back)
print("handoff: arrive->mma-wakes", med(diff(P2, M1)[lo:hi]), "  mma-issued->producer-wakes(next use of the stage)",
      med([(P1[i + 2] - M2[i]) if (P1[i + 2] and M2[i]) else None for i in range(lo, hi)]))
