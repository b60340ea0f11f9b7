"""Diagnostic (GPU): probe the tcgen05 GEMM on chosen shapes and say what the output looks like."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import freeimpala_b200 as fi

def run(trans, m, n, k, seed=0, ones=False):
    rng = np.random.default_rng(seed)
    shp_a = (k, m) if trans == "TN" else (m, k)
    shp_b = (n, k) if trans == "NT" else (k, n)
    A = np.ones(shp_a, np.float32) if ones else rng.standard_normal(shp_a).astype(np.float32)
    B = np.ones(shp_b, np.float32) if ones else rng.standard_normal(shp_b).astype(np.float32)
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    opA = A64.T if trans == "TN" else A64
    opB = B64.T if trans == "NT" else B64
    ref = opA @ opB
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dC = torch.full((m, n), float("nan"), device="cuda")
    wsb = fi.ops.gemm_workspace_bytes(trans, m, n, k, "tcgen05")
    ws = torch.zeros(max(wsb, 16), dtype=torch.uint8, device="cuda")
    try:
        fi.ops.gemm(trans, m, n, k, dA.data_ptr(), A.shape[1], dB.data_ptr(), B.shape[1], dC.data_ptr(), n, None, False,
                    "tcgen05", ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"{trans} m={m} n={n} k={k}: EXC {e}")
        return
    got = dC.cpu().numpy().astype(np.float64)
    err = np.abs(got - ref).max()
    msg = f"{trans} m={m} n={n} k={k} ones={ones}: maxerr {err:.3e} nan {np.isnan(got).mean():.2f} zero {(got == 0).mean():.2f}"
    # does the output match a prefix / subset of k-blocks?
    nkb = (k + 31) // 32
    for j in range(1, nkb + 1):
        part = opA[:, :32 * j] @ opB[:32 * j, :]
        if np.abs(got - part).max() < 1e-3 * max(1, np.abs(part).max()):
            msg += f" == first {j}/{nkb} k-blocks"
            break
    for j in range(nkb):
        part = opA[:, 32 * j:32 * j + 32] @ opB[32 * j:32 * j + 32, :]
        if np.abs(got - part).max() < 1e-3 * max(1, np.abs(part).max()):
            msg += f" == only k-block {j}"
    print(msg, "sample", got[0, :4], "ref", ref[0, :4])

for k in (32, 64, 96, 128, 160, 512):
    run("NT", 128, 256, k)
run("NT", 128, 32, 512)
run("NT", 128, 128, 512)
run("NT", 128, 128, 512, ones=True)
for k in (8, 32, 64):
    run("NN", 128, 256, k)
    run("NN", 128, 256, k, ones=True)
    run("TN", 128, 256, k)
    run("TN", 128, 256, k, ones=True)
run("NN", 128, 32, 32)
run("TN", 128, 32, 32)
