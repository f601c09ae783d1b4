"""%globaltimer trace of one cluster of the 4-CTA InfoNCE backward (infonce_bwd_tc5.cu built with -DDMF_TC5_TRACE,
run with DMF_BWD_TC5=1): when each pair issued S(k), saw its own / the foreign W tile, and how long the softmax took."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from disentagled_multimodal_fusion_b200 import ops  # noqa: E402
from disentagled_multimodal_fusion_b200._lib import lib, check, ptr, stream  # noqa: E402
B, D = 65536, 512
dev = "cuda"
torch.manual_seed(0)
z0 = torch.nn.functional.normalize(torch.randn(B, D, device=dev), dim=-1)
z1 = torch.nn.functional.normalize(0.5 * z0 + 0.1 * torch.randn(B, D, device=dev), dim=-1)
b0, b1 = ops.cast_bf16(z0), ops.cast_bf16(z1)
lse = torch.full((2, B), 12.0, device=dev)
b1T = ops.transpose_bf16(b1)
dz = torch.empty(B, D, device=dev)
one = torch.ones(1, device=dev)
scale = 1 / 0.07
for _ in range(3):
    check(lib.dmf_infonce_bwd(ptr(b0), D, B, ptr(lse[0]), ptr(b1), D, ptr(b1T), b1T.stride(0), B, ptr(lse[1]), D, scale, scale / (2 * B), ptr(one), 0, ptr(dz), D, 0, 1, stream()))
torch.cuda.synchronize()
out = np.zeros((8, 2048), dtype=np.uint64)
f = lib.dmf_tc5_trace_read
f.argtypes = [ctypes.c_void_p]; f.restype = ctypes.c_int
print("rc", f(out.ctypes.data))
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/tc5_trace.npy", out)
t0 = out[out > 0].min()
rel = (out.astype(np.int64) - int(t0))
names = ["p0 S issued", "p0 own W ready (p_full)", "p0 foreign W ready", "p0 softmax s_full/done", "p1 S issued", "p1 own W ready", "p1 foreign W ready", "p1 softmax"]
for i in range(8):
    print(names[i], rel[i, 100:112].tolist())
