"""Per-shape timing of the tower GEMMs (rh_gemm_tf32x3): forward, input-gradient and weight-gradient products of the
429-256-128 tower at batch 4096, each replayed from a CUDA graph and timed with CUDA events; the weight-gradient products are
swept over split-K.  Prints one CSV line per (shape, split_k): us per launch, CTAs, TF/s (fp32-equivalent flops).

    python tools/sweep_gemm.py > gpurun_out/gemm_sweep.csv
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-rechub_b200"))
import torch  # noqa: E402

from torch_rechub.b200 import ops  # noqa: E402

DEV = "cuda:0"
ROWS = int(os.environ.get("SWEEP_ROWS", "4096"))


def time_graph(fn, reps=50):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def main():
    torch.manual_seed(0)
    print("name,M,N,K,a_mn,b_mn,split_k,ctas,us,tflops_fp32_equiv")
    pad = ops._pad4
    for name, k_in, cols in (("L1", 429, 256), ("L2", 256, 128)):
        x = torch.randn(ROWS, pad(k_in), device=DEV)[:, :k_in]
        w = torch.randn(cols, pad(k_in), device=DEV)[:, :k_in]
        dh = torch.randn(ROWS, cols, device=DEV)
        bias = torch.randn(cols, device=DEV)
        h = torch.empty(ROWS, cols, device=DEV)
        dx = torch.empty(ROWS, pad(k_in), device=DEV)
        cases = [("%s_fwd" % name, lambda: ops.gemm3x(x, False, w, False, ROWS, cols, k_in, bias=bias, out=h), ROWS, cols, k_in, 0, 0, 1),
                 ("%s_dX" % name, lambda: ops.gemm3x(dh, False, w, True, ROWS, k_in, cols, out=dx), ROWS, k_in, cols, 0, 1, 1)]
        for split in (1, 2, 4, 8, 16, 32, 64, 128):
            dw = torch.zeros(cols, k_in, device=DEV)
            cases.append(("%s_dW" % name, (lambda s=split, o=dw: ops.gemm3x(dh, True, x, True, cols, k_in, ROWS, split_k=s, out=o)), cols, k_in, ROWS, 1, 1, split))
        for label, fn, M, N, K, amn, bmn, split in cases:
            us = time_graph(fn)
            kb = (K + 31) // 32
            eff_split = min(split, kb)
            ctas = ((M + 127) // 128) * ((N + 127) // 128) * eff_split
            print("%s,%d,%d,%d,%d,%d,%d,%d,%.2f,%.1f" % (label, M, N, K, amn, bmn, split, ctas, us, 2.0 * M * N * K / us / 1e6))


if __name__ == "__main__":
    main()
