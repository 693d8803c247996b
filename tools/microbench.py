"""Graph-timed microbenchmarks of the decode-step kernels (GPU box).  python tools/microbench.py"""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def child():
    import whisperkit_b200 as wk
    from whisperkit_b200._lib import check
    B = 64
    m = wk.Model("large-v3", max_batch=B); m.init_random(1)
    dec = wk.TextDecoder(m, B)
    names = {0: "cross_attn", 4: "dec_qkv_gemm", 6: "dec_o_gemm", 7: "dec_fc2_gemm", 8: "reduce_resid_ln", 9: "self_attn@100"}
    f, w = C.c_float(), C.c_double()
    out = []
    for k, n in names.items():
        check(m.lib.wk_bench_kernel(m.handle, dec.handle, k, B, 200, C.byref(f), C.byref(w)))
        out.append(f"{n}={f.value*1000:.2f}us")
    print("  " + "  ".join(out), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(); sys.exit(0)
    variants = [("stream-launch baseline", {}), ("graph", {"WKB200_BENCH_GRAPH": "1"}),
                ("graph tmem512", {"WKB200_BENCH_GRAPH": "1", "WKB200_GEMM_TMEM": "512"}),
                ("graph stages2", {"WKB200_BENCH_GRAPH": "1", "WKB200_GEMM_STAGES": "2"}),
                ("graph exit-after-setup", {"WKB200_BENCH_GRAPH": "1", "WKB200_GEMM_DEBUG": "1"}),
                ("graph no-epilogue-stores", {"WKB200_BENCH_GRAPH": "1", "WKB200_GEMM_DEBUG": "2"}),
                ("graph pdl", {"WKB200_BENCH_GRAPH": "1", "WKB200_PDL": "1"})]
    for name, env in variants:
        print("==", name, flush=True)
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, "child"], env=e, timeout=200)
