"""GEMM-view probe of the contraction kernel (development): prints TFLOP/s at 8192^3 for the current env knobs."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import aces4_b200 as sip
api = sip.api
sip.init(0)
stream = torch.cuda.ExternalStream(api.stream_handle())
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
A, B, Cc = api.DeviceBlock((n, n)).fill(0.5), api.DeviceBlock((n, n)).fill(0.25), api.DeviceBlock((n, n))
for _ in range(2):
    api.dgemm_tn(n, n, n, A, n, B, n, Cc, n)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sip.sync()
e0.record(stream)
for _ in range(4):
    api.dgemm_tn(n, n, n, A, n, B, n, Cc, n)
e1.record(stream)
e1.synchronize()
print(json.dumps({"dbg": os.environ.get("SIPGPU_DBG", "0"), "w16": os.environ.get("SIPGPU_WARPS16", "0"), "n": n,
                  "tflops": round(2.0 * n ** 3 / (e0.elapsed_time(e1) / 4) / 1e9, 2)}))
