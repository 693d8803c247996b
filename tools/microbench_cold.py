"""Cold-weight (HBM-streamed) timings of the decoder swap-AB GEMMs versus split-K depth.  python tools/microbench_cold.py"""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def child():
    import whisperkit_b200 as wk
    from whisperkit_b200._lib import check
    B = 64
    m = wk.Model("large-v3", max_batch=B); m.init_random(1)
    dec = wk.TextDecoder(m, B)
    names = {14: "dxd(o/cq/co)", 17: "qkv", 15: "fc1", 16: "fc2", 8: "reduce_ln"}
    f, w = C.c_float(), C.c_double()
    out = []
    for k, n in names.items():
        check(m.lib.wk_bench_kernel(m.handle, dec.handle, k, B, 192, C.byref(f), C.byref(w)))
        out.append(f"{n}={f.value*1000:.2f}us")
    print("  " + "  ".join(out), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(); sys.exit(0)
    for s in (0, 1, 2, 4, 5, 10, 20):
        print("== forced splits", s or "default", flush=True)
        e = dict(os.environ); e["WKB200_BENCH_GRAPH"] = "1"
        if s:
            e["WKB200_FORCE_SPLITS"] = str(s)
        subprocess.run([sys.executable, __file__, "child"], env=e, timeout=200)
