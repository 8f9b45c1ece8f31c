"""Timeline of CTA 0 of the LSTM tcgen05 step kernel (one layer, last timestep): where an item's time goes, per role.
usage: python scratch/lstm_trace_digest.py [layer ...]   (env TR_M machines, TR_ROWS rows, TR_L lookback)"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import __graft_entry__ as ge
ge.build()
from gordo_components_b200 import engine, _cabi
from benchmarks import secondary as sec
lib = _cabi.load_library()
CAP = 2048
M, ROWS, LB = int(os.environ.get('TR_M', 32)), int(os.environ.get('TR_ROWS', 10000)), int(os.environ.get('TR_L', 6))
layers = [int(v) for v in sys.argv[1:]] or [0, 1, 3, 5]
sec.lstm_share(torch, engine, machines=M, rows=ROWS, lookback=LB)  # warm-up, workspaces
for layer in layers:
    buf = torch.zeros(3 + 3 * CAP, dtype=torch.int64, device='cuda')
    lib.gb_debug_set_lstm_trace(C.c_void_p(buf.data_ptr()), CAP, layer)
    sec.lstm_share(torch, engine, machines=M, rows=ROWS, lookback=LB)
    torch.cuda.synchronize()
    lib.gb_debug_set_lstm_trace(None, 0, -1)
    b = buf.cpu().numpy()
    ev = {}
    for role in range(3):
        for v in b[3 + role * CAP: 3 + role * CAP + int(b[role])]:
            v = int(v) & ((1 << 64) - 1)
            ev[(v & 0xf, (v >> 8) & 0xfff, (v >> 4) & 0xf)] = v >> 20
    items = sorted({n for (c, n, k) in ev if c == 6})
    nch = 1 + max(k for (c, n, k) in ev if c == 4)
    # the recorder thread's epilogue group takes every item (layer 0) or every other item (two groups): its period spans `st` items
    st = 2 if len(items) > 3 and all(n % 2 == 0 for n in items) else 1
    steady = [n for n in items[3:-1] if (6, n + st, 0) in ev and (10, n, 0) in ev and (3, n, 0) in ev and (5, n, nch - 1) in ev]
    def mean(f):
        vals = [f(n) for n in steady]
        return float(np.mean(vals)) if vals else float('nan')
    print(f"layer {layer}: {len(items)} items on CTA 0, {nch} chunks per item, {len(steady)} in steady state; cycles per item (mean)")
    print(f"  epilogue thread 0 ({st} group{'s' if st > 1 else ''}): period per item {mean(lambda n: (ev[(6, n + st, 0)] - ev[(6, n, 0)]) / st):8.0f}; of its own item: prologue (state / bias requests) {mean(lambda n: ev[(7, n, 0)] - ev[(6, n, 0)]):7.0f}"
          f" + wait for the accumulator {mean(lambda n: ev[(8, n, 0)] - ev[(7, n, 0)]):7.0f} + gates/cell {mean(lambda n: ev[(9, n, 0)] - ev[(8, n, 0)]):7.0f}"
          f" + h stores {mean(lambda n: ev[(10, n, 0)] - ev[(9, n, 0)]):7.0f} + to its next item {mean(lambda n: ev[(6, n + st, 0)] - ev[(10, n, 0)]):7.0f}")
    print(f"  MMA issuer: wait for a free accumulator {mean(lambda n: ev[(3, n, 0)] - ev[(5, n - 1, nch - 1)] if (5, n - 1, nch - 1) in ev else 0):7.0f};"
          f" per chunk: wait for the stage {mean(lambda n: np.mean([ev[(4, n, k)] - (ev[(5, n, k - 1)] if k else ev[(3, n, 0)]) for k in range(nch)])):7.0f},"
          f" issue + commit {mean(lambda n: np.mean([ev[(5, n, k)] - ev[(4, n, k)] for k in range(nch)])):6.0f};"
          f" first wait to last commit {mean(lambda n: ev[(5, n, nch - 1)] - ev[(3, n, 0)]):8.0f}")
    pk = [n for n in steady if all((1, n, k) in ev and (2, n, k) in ev for k in range(nch))]
    if pk:
        print(f"  TMA producer: per chunk wait for an empty stage {np.mean([ev[(1, n, k)] - (ev[(2, n, k - 1)] if k else ev[(2, n - 1, nch - 1)]) for n in pk for k in range(nch) if k or (2, n - 1, nch - 1) in ev]):7.0f},"
              f" issue {np.mean([ev[(2, n, k)] - ev[(1, n, k)] for n in pk for k in range(nch)]):6.0f};"
              f" stage issued -> stage landed (MMA side saw it) {np.mean([ev[(4, n, k)] - ev[(2, n, k)] for n in pk for k in range(nch) if (4, n, k) in ev]):7.0f}")
