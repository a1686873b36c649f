"""Developer tool: brta_pindicator_counts against the HBM roofline.

    python tools/perf_pindicator.py [S] [N] [K] [C]
Algorithmic bytes per launch = S*N (labels in) + 4*N*C (counts out) + S*K (table)."""
import ctypes as C
import json
import os
import sys


sys.path.insert(0, '.')
import torch  # noqa: E402
from basicrta_b200 import _cabi  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 15
NC = int(sys.argv[4]) if len(sys.argv) > 4 else 4
lib = _cabi.load()
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(1)
ind = torch.randint(0, K, (S, N), dtype=torch.uint8, device=dev, generator=g)
lut = torch.randint(-1, NC, (S, K), dtype=torch.int8, device=dev, generator=g)
counts = torch.zeros((N, NC), dtype=torch.int32, device=dev)
stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    rc = lib.brta_pindicator_counts(C.c_void_p(ind.data_ptr()), N, S, N, C.c_void_p(lut.data_ptr()), K, NC,
                                    C.c_void_p(counts.data_ptr()), stream)
    assert rc == 0, lib.brta_last_error()


for _ in range(3):
    run()
best = 1e30
for _ in range(5):
    counts.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
nbytes = S * N + 4 * N * NC + S * K
peak = 6546.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
print(('generic ' if os.environ.get('BRTA_PINDICATOR_GENERIC') else 'class   ') + f'S={S} N={N} K={K} C={NC}: {best:.3f} ms, {nbytes / best / 1e6:.1f} GB/s algorithmic '
      f'({nbytes / best / 1e6 / peak:.3f} of {peak:.0f} GB/s), {S * N / best / 1e6:.1f} G labels/s')
