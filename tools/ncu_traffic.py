"""Turns the ncu --set full reports of tools/gpu_profile.sh into the tracked evidence under profiles/:

    python tools/ncu_traffic.py gpurun_out/prof r02

writes profiles/r02_traffic.json (per kernel: DRAM bytes read + written per launch, duration, tensor-pipe / DRAM utilisation as ncu reports
them) and profiles/r02_ncu_summary.md, and copies the .ncu-rep files to profiles/r02_<name>.ncu-rep.  Needs only `ncu -i` (no GPU)."""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__grid_size", "launch__block_size"]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}.get(unit, 1)


def read_report(path):
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(r.stdout)))
    if len(rows) < 3:
        return []
    head, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = {}
        for h, u, v in zip(head, units, row):
            d[h] = (v, u)
        out.append(d)
    return out


def main():
    src, tag = sys.argv[1], sys.argv[2]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prof = os.path.join(root, "profiles")
    traffic, lines = {}, [f"# ncu --set full captures ({tag})", "", "| report | kernel | grid x block | duration us | DRAM read MB | DRAM write MB | DRAM % | tensor pipe % | regs |", "|---|---|---|---|---|---|---|---|---|"]
    for fn in sorted(os.listdir(src)):
        if not fn.endswith(".ncu-rep"):
            continue
        shutil.copy(os.path.join(src, fn), os.path.join(prof, f"{tag}_{fn}"))
        for i, d in enumerate(read_report(os.path.join(src, fn))):
            name = d.get("Kernel Name", ("?", ""))[0].replace("(anonymous namespace)::", "")
            def g(key, conv=None):
                if key not in d:
                    return None
                v, u = d[key]
                try:
                    return conv(v, u) if conv else float(v.replace(",", ""))
                except ValueError:
                    return None
            dur = g("gpu__time_duration.sum", to_us)
            rd, wr = g("dram__bytes_read.sum", to_bytes), g("dram__bytes_write.sum", to_bytes)
            tens = g("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active") or g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
            dram = g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
            key = f"{fn[:-8]}#{i}:{name.split('(')[0]}"
            traffic[key] = {"kernel": name, "report": f"{tag}_{fn}", "launch": i, "duration_us": dur, "dram_read_bytes": rd, "dram_write_bytes": wr,
                            "bytes": (rd or 0) + (wr or 0), "dram_pct": dram, "tensor_pipe_pct": tens,
                            "grid": g("launch__grid_size"), "block": g("launch__block_size"), "registers": g("launch__registers_per_thread"),
                            "inst_executed": g("smsp__inst_executed.sum"), "smem_bank_conflicts": g("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")}
            lines.append(f"| {tag}_{fn} | `{name.split('(')[0][:60]}` | {g('launch__grid_size')} x {g('launch__block_size')} | {dur and round(dur, 1)} | "
                         f"{rd and round(rd / 1e6, 1)} | {wr and round(wr / 1e6, 1)} | {dram} | {tens} | {g('launch__registers_per_thread')} |")
    # bench.py looks its kernel table up by name prefix: which capture (report # launch) stands for which table entry
    roles = {"decoder_cross_attention_kernel": ["cross_attention#0"], "gemm_tcgen05_kernel[enc QKV": ["encoder_gemm#0"],
             "gemm_tcgen05_kernel[enc out-proj": ["encoder_gemm#1"], "gemm_tcgen05_kernel[enc FC1+GELU": ["encoder_gemm#2"],
             "gemm_tcgen05_kernel[enc FC2": ["encoder_gemm#3"], "encoder_attention_tcgen05_kernel": ["encoder_attention#0"],
             "mel_pass1+pass2": ["mel#0", "mel#1"], "decoder_self_attention_kernel": ["self_attention#0"], "sampler_kernel": ["sampler#0"]}
    by_prefix = {}
    for role, caps in roles.items():
        ents = [v for c in caps for k, v in traffic.items() if k.startswith(c + ":")]
        if len(ents) == len(caps):
            by_prefix[role] = {"bytes": sum(e["bytes"] for e in ents), "duration_us": sum(e["duration_us"] or 0 for e in ents),
                               "tensor_pipe_pct": ents[0]["tensor_pipe_pct"], "dram_pct": ents[0]["dram_pct"], "captures": caps,
                               "config": "whisper-large-v3, 64 windows (bench.py --profile-pass), ncu --set full --clock-control none"}
    json.dump({"by_bench_kernel_prefix": by_prefix, "captures": traffic}, open(os.path.join(prof, f"{tag}_traffic.json"), "w"), indent=1)
    # SASS evidence from the built library
    lib = os.path.join(root, "whisperkit_b200", "libwkb200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    lines += ["", "## SASS mnemonics per kernel (cuobjdump -sass libwkb200.so)", "", "| kernel | UTC*MMA | UTMALDG | UBLKCP | LDTM | HMMA |", "|---|---|---|---|---|---|"]
    cur, counts = None, {}
    for ln in sass.splitlines():
        if "Function :" in ln:
            cur = ln.split("Function :")[1].strip()
            counts[cur] = [0, 0, 0, 0, 0]
        elif cur:
            for j, m in enumerate(("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", " HMMA")):
                if m in ln:
                    counts[cur][j] += 1
    for k, c in counts.items():
        if any(c):
            d = (subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip() or k).replace("(anonymous namespace)::", "")
            lines.append(f"| `{d.split('(')[0][:70]}` | {c[0]} | {c[1]} | {c[2]} | {c[3]} | {c[4]} |")
    open(os.path.join(prof, f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
