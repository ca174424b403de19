#!/usr/bin/env python
"""Turn the raw ncu outputs under gpurun_out/ into the tracked summaries under profiles/.

    python tools/ncu_summary.py <tag>        # e.g. r01  -> profiles/r01_launches.md, r01_gemm.md, r01_attention.md,
                                             #              r01_gemm_traffic.json

Inputs (produced on the GPU box, see profiles/README.md for the exact commands):
    gpurun_out/<tag>_launches.csv      ncu --metrics gpu__time_duration.sum launch list of bench.py
    gpurun_out/<tag>_prof_gemm.ncu-rep ncu --set full capture of gemm_tc_kernel launches
    gpurun_out/<tag>_prof_attn.ncu-rep ncu --set full capture of attention_tc_kernel launches
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def to_bytes(val, unit):
    v = float(val)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def launches(tag):
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = {n: i for i, n in enumerate(rows[hi])}
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        name = re.sub(r"\(CUtensorMap.*|\(const.*|\(float.*", "", r[h["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "")
        v = float(r[h["Metric Value"]]) / {"ns": 1000.0, "us": 1.0, "ms": 1e-3}.get(r[h["Metric Unit"]], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: launch list of `bench.py --steps 2 --warmup 3` (ncu, cold cache, serialised: compare SHARES)\n\n")
        f.write(f"{n} launches captured (≈2.2 DiT steps), {tot / 1e3:.2f} ms summed device time.\n\n| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |\n")
    return agg, tot


def raw_table(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def kernel_report(tag, what, title):
    rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof_{what}.ncu-rep")
    if not os.path.exists(rep):
        return None
    hdr, units, rows = raw_table(rep)
    h = {n: i for i, n in enumerate(hdr)}
    recs = []
    with open(os.path.join(OUT, f"{tag}_{ {'gemm': 'gemm', 'attn': 'attention', 'vae': 'vae'}[what] }.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none` — {title}\n\nPer launch (cold cache, ~40 replays; not a timing source):\n\n")
        for r in rows:
            name = re.sub(r"\(CUtensorMap.*", "", r[h["Kernel Name"]]).replace("void ", "")
            f.write(f"## `{name[:100]}`\n\n| metric | value |\n|---|---|\n")
            rec = {"kernel": name}
            for k in KEYS:
                if k in h:
                    f.write(f"| {k} | {r[h[k]]} {units[h[k]]} |\n")
                    rec[k] = (r[h[k]], units[h[k]])
            f.write("\n")
            recs.append(rec)
        # instruction-level hot spots of the first kernel
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        if len(srows) > 3:
            shdr = srows[1]
            sh = {n: i for i, n in enumerate(shdr)}
            body = []
            for r in srows[2:]:
                if r and r[0] == "Kernel Name":
                    break
                if len(r) == len(shdr):
                    body.append(r)
            tot = sum(int(r[sh["# Samples"]]) for r in body) or 1
            f.write("## Top stall sites of the first captured launch (warp-state sampling)\n\n| samples | share | SASS | top stall reasons |\n|---:|---:|---|---|\n")
            for r in sorted(body, key=lambda r: -int(r[sh["# Samples"]]))[:14]:
                st = {n: int(r[sh[n]]) for n in shdr if n.startswith("stall_") and "(Not" not in n and int(r[sh[n]]) > 0}
                top = ", ".join(f"{k[6:]} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
                f.write(f"| {r[sh['# Samples']]} | {100 * int(r[sh['# Samples']]) / tot:.1f}% | `{r[sh['Source']].strip()[:70]}` | {top} |\n")
            mn = collections.Counter()
            for r in body:
                for m in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "MUFU.EX2", "HMMA"):
                    if m in r[sh["Source"]]:
                        mn[m] += 1
            f.write("\nBlackwell mnemonics present in the captured SASS: " + ", ".join(f"{k}×{v}" for k, v in mn.items()) + "\n")
    return recs


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    agg, tot = launches(tag)
    g = kernel_report(tag, "gemm", "gemm_tc_kernel launches of one DiT block (to_q, proj, qkv, proj, fc1, fc2)")
    kernel_report(tag, "attn", "attention_tc_p_kernel (cross- then self-attention of one block)")
    kernel_report(tag, "vae", "VAE decoder: conv3_halo_kernel / groupnorm_silu_kernel launches of the 8^3 tail (tools/vae_decode_perf.py)")
    if g:
        tr = [to_bytes(*r["dram__bytes_read.sum"]) + to_bytes(*r["dram__bytes_write.sum"]) for r in g if "dram__bytes_read.sum" in r]
        json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "launches": len(tr), "per_launch": tr,
                   "source": f"profiles/{tag}_gemm.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
                  open(os.path.join(OUT, f"{tag}_gemm_traffic.json"), "w"), indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
