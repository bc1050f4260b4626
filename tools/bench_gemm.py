"""Micro-benchmark of lav_gemm_f16 on the shapes of the hot path (CUDA events, L2 flushed between runs)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lavender_b200 import ops, _lib as L  # noqa: E402


def timeit(fn, iters=10, flush=None):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    torch.manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    rows = []
    shapes = [  # (tag, M, N, K, mode)
        ("swin_s0_qkv", 125440, 384, 128, "fwd"), ("swin_s0_fc1_gelu", 125440, 512, 128, "gelu"),
        ("swin_s0_fc2_res", 125440, 128, 512, "res"), ("swin_s2_qkv", 7840, 1536, 512, "fwd"),
        ("swin_s2_fc1_gelu", 7840, 2048, 512, "gelu"), ("swin_s2_fc2_res", 7840, 512, 2048, "res"),
        ("bert_qkv_vtm", 9056, 2304, 768, "fwd"), ("bert_ffn1_vtm", 9056, 3072, 768, "gelu"),
        ("bert_ffn2_vtm", 9056, 768, 3072, "fwd32"), ("mlm_decoder", 1352, 30522, 768, "fwd32"),
        ("square_8k", 8192, 8192, 8192, "fwd"),
        ("swin_s2_fc1_dgrad", 7840, 512, 2048, "dgrad"), ("swin_s2_fc1_wgrad", 2048, 512, 7840, "wgrad"),
        ("swin_s0_qkv_wgrad", 384, 128, 125440, "wgrad"), ("bert_ffn1_wgrad", 3072, 768, 9056, "wgrad"),
    ]
    shapes += [("swin_s2_fc2_dgrad_gelu", 7840, 2048, 512, "dgrad_gelu"), ("swin_s2_proj_res", 7840, 512, 512, "res")]
    if "--sweep" in sys.argv:
        shapes = [(f"k{K}", 7840, 2048, K, "fwd") for K in (64, 256, 512, 1024, 2048, 4096)]
        shapes += [(f"m{M}", M, 2048, 512, "fwd") for M in (128, 1024, 18944, 37888)]
        shapes += [("swin_s2_fc1_gelu", 7840, 2048, 512, "gelu"), ("swin_s2_fc2_res", 7840, 512, 2048, "res"),
                   ("swin_s0_fc1_gelu", 125440, 512, 128, "gelu"), ("swin_s2_fc1_wgrad", 2048, 512, 7840, "wgrad")]
    if "--hot" in sys.argv:   # every GEMM shape of a Swin stage-2 block and of a BERT layer at B = 8 (configs[1])
        shapes = [("s2_qkv", 7840, 1536, 512, "fwd"), ("s2_proj_res", 7840, 512, 512, "res"),
                  ("s2_fc1_gelu", 7840, 2048, 512, "gelu"), ("s2_fc2_res", 7840, 512, 2048, "res"),
                  ("s2_fc2_dgrad_gelu", 7840, 2048, 512, "dgrad_gelu"), ("s2_fc1_dgrad", 7840, 512, 2048, "dgrad"),
                  ("s2_proj_dgrad", 7840, 512, 512, "dgrad"), ("s2_qkv_dgrad", 7840, 512, 1536, "dgrad"),
                  ("s2_fc1_wgrad", 2048, 512, 7840, "wgrad"), ("s2_fc2_wgrad", 512, 2048, 7840, "wgrad"),
                  ("s2_proj_wgrad", 512, 512, 7840, "wgrad"), ("s2_qkv_wgrad", 1536, 512, 7840, "wgrad"),
                  ("s1_qkv", 31360, 768, 256, "fwd"), ("s1_fc1_gelu", 31360, 1024, 256, "gelu"),
                  ("s1_fc2_res", 31360, 256, 1024, "res"), ("s1_fc2_dgrad_gelu", 31360, 1024, 256, "dgrad_gelu"),
                  ("bert_qkv", 11360, 2304, 768, "fwd"), ("bert_out", 11360, 768, 768, "fwd32"),
                  ("bert_ffn1", 11360, 3072, 768, "gelu"), ("bert_ffn2", 11360, 768, 3072, "fwd32"),
                  ("bert_ffn2_dgrad_gelu", 11360, 3072, 768, "dgrad_gelu"), ("bert_ffn1_dgrad", 11360, 768, 3072, "dgrad"),
                  ("bert_qkv_dgrad", 11360, 768, 2304, "dgrad"), ("bert_ffn1_wgrad", 3072, 768, 11360, "wgrad"),
                  ("bert_qkv_wgrad", 2304, 768, 11360, "wgrad")]
    if "--head" in sys.argv:  # MLM head (vocab 30522) on the labelled rows: 128 MLM + 32 VTM, separately and merged
        shapes = [("dec_fwd_128", 128, 30522, 768, "fwd32"), ("dec_fwd_32", 32, 30522, 768, "fwd32"),
                  ("dec_fwd_160", 160, 30522, 768, "fwd32"),
                  ("dec_dgrad_128", 128, 768, 30528, "dgrad"), ("dec_dgrad_32", 32, 768, 30528, "dgrad"),
                  ("dec_dgrad_acc_128", 128, 768, 30528, "dgrad_acc"), ("dec_dgrad_acc_160", 160, 768, 30528, "dgrad_acc"),
                  ("dec_wgrad_128", 30528, 768, 128, "wgrad"), ("dec_wgrad_32", 30528, 768, 32, "wgrad"),
                  ("dec_wgrad_160", 30528, 768, 160, "wgrad")]   # (vocab padded to 30528: 16-byte aligned fp16 rows)
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:
        shapes = [s_ for s_ in shapes if s_[0] in only]
    for tag, M, N, K, mode in shapes:
        if mode == "wgrad":
            a = torch.randn(K, M, device="cuda").half()
            b = torch.randn(K, N, device="cuda").half()
            out = torch.zeros(M, N, device="cuda")
            bg = torch.zeros(M, device="cuda") if ("--hot" in sys.argv or "--head" in sys.argv) else None   # the path's wgrads carry the bias gradient
            fn = lambda: ops.gemm(a, b, out, M=M, N=N, K=K, a_major=1, b_major=1, accumulate=True, bias_grad=bg)
        elif mode == "dgrad_acc":   # split-K dgrad into a zeroed fp32 buffer (skinny M, huge K)
            a = torch.randn(M, K, device="cuda").half()
            b = torch.randn(K, N, device="cuda").half()
            out = torch.zeros(M, N, device="cuda")
            fn = lambda: ops.gemm(a, b, out, M=M, N=N, K=K, b_major=1, accumulate=True)
        elif mode in ("dgrad", "dgrad_gelu"):
            a = torch.randn(M, K, device="cuda").half()
            b = torch.randn(K, N, device="cuda").half()
            out = torch.zeros(M, N, device="cuda").half()
            if mode == "dgrad":
                fn = lambda: ops.gemm(a, b, out, M=M, N=N, K=K, b_major=1)
            else:
                aux = torch.randn(M, N, device="cuda").half()
                fn = lambda: ops.gemm(a, b, out, M=M, N=N, K=K, b_major=1, act=L.ACT_GELU_BWD, aux=aux)
        else:
            a = torch.randn(M, K, device="cuda").half()
            b = torch.randn(N, K, device="cuda").half()
            bias = torch.randn(N, device="cuda")
            if mode == "fwd":
                out = torch.zeros(M, N, device="cuda").half()
                fn = lambda: ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias)
            elif mode == "fwd32":
                out = torch.zeros(M, N, device="cuda")
                fn = lambda: ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias)
            elif mode == "gelu":
                out = torch.zeros(M, N, device="cuda").half()
                aux = torch.zeros(M, N, device="cuda").half()
                fn = lambda: ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias, act=L.ACT_GELU, aux=aux)
            else:
                out = torch.zeros(M, N, device="cuda")
                res = torch.randn(M, N, device="cuda")
                fn = lambda: ops.gemm(a, b, out, M=M, N=N, K=K, bias=bias, residual=res)
        ms = timeit(fn, flush=flush)
        tf = 2.0 * M * N * K / ms / 1e9
        # torch (cuBLAS) for comparison on the plain product
        if "--no-cublas" in sys.argv:
            ms_t = float("nan")
        elif mode == "wgrad":
            ms_t = timeit(lambda: torch.matmul(a.t(), b), flush=flush)
        elif mode in ("dgrad", "dgrad_gelu", "dgrad_acc"):
            ms_t = timeit(lambda: torch.matmul(a, b), flush=flush)
        else:
            ms_t = timeit(lambda: torch.matmul(a, b.t()), flush=flush)
        rows.append(dict(tag=tag, M=M, N=N, K=K, mode=mode, ms=round(ms, 4), tflops=round(tf, 1),
                         cublas_ms=round(ms_t, 4), cublas_tflops=round(2.0 * M * N * K / ms_t / 1e9, 1)))
        print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    name = os.environ.get("LAV_BENCH_GEMM_OUT", "bench_gemm.json")
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", name), "w"), indent=1)


if __name__ == "__main__":
    main()
