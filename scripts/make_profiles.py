#!/usr/bin/env python
"""Turn the raw artefacts of a GPU batch (gpurun_out/*.ncu-rep) and the built extension into the
reviewable files under profiles/r2/: ncu summaries, the full SASS of the flagship kernel, a SASS
mnemonic inventory of every kernel, and the PTX lines that prove multimem / TMA use.
Runs on the CPU box (ncu -i, cuobjdump)."""
import collections
import csv
import io
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles" / "r2"
OUT.mkdir(parents=True, exist_ok=True)
SO = next((ROOT / "torch_cgx_b200").glob("_C*.so"))

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "lts__t_sector_hit_rate.pct",
]
STALLS = ["long_scoreboard", "wait", "short_scoreboard", "barrier", "not_selected", "no_instruction",
          "branch_resolving", "math_pipe_throttle", "lg_throttle", "mio_throttle", "drain", "membar"]


def run(cmd):
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def ncu_summary(rep: Path, title: str, note: str):
    raw = run(["ncu", "-i", str(rep), "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        return
    hdr = rows[0]
    lines = [f"# {title}", "", note, "", f"source: `gpurun_out/{rep.name}` (`ncu --set full --clock-control none --import-source on`)", ""]
    for r in rows[2:]:
        lines += [f"## `{r[hdr.index('Kernel Name')]}`", "", "| metric | value |", "|---|---|"]
        for m in METRICS:
            if m in hdr:
                lines.append(f"| {m} | {r[hdr.index(m)]} |")
        for s in STALLS:
            m = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if m in hdr:
                lines.append(f"| stall {s} (warps per issue) | {float(r[hdr.index(m)]):.3f} |")
        lines.append("")
    src = run(["ncu", "-i", str(rep), "--page", "source", "--csv"])
    srows = list(csv.reader(io.StringIO(src)))
    if len(srows) > 3:
        h = srows[1]
        try:
            i_src, i_ex, i_st = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
            data = []
            for r in srows[2:]:
                if len(r) <= max(i_src, i_ex, i_st):
                    if data:
                        break  # next kernel's header: only the first kernel is listed
                    continue
                try:
                    data.append((r[i_src], int(r[i_ex] or 0), int(r[i_st] or 0)))
                except ValueError:
                    if data:
                        break
            total = sum(d[2] for d in data) or 1
            lines += ["## hottest instructions of the first kernel (warp-stall samples)", "", "| samples | % | executed | SASS |", "|---|---|---|---|"]
            for d in sorted(data, key=lambda d: -d[2])[:15]:
                lines.append(f"| {d[2]} | {100 * d[2] / total:.1f} | {d[1]} | `{d[0][:90]}` |")
            lines.append("")
        except ValueError:
            pass
    (OUT / f"ncu_{rep.stem.replace('a_prof_', '').replace('f_prof_', '')}.md").write_text("\n".join(lines))


def sass():
    txt = run(["cuobjdump", "-sass", str(SO)])
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    inv = ["# SASS inventory of the shipped extension (`cuobjdump -sass torch_cgx_b200/_C*.so`, sm_100a)", "",
           "Mnemonics that matter for the review: `UBLKCP` = TMA bulk copy (`cp.async.bulk`), `SYNCS` = mbarrier,",
           "`LDGMC` = `multimem.ld_reduce` (NVLS in-switch reduction), `LDG.E.ENL2.256` / `STG.E.ENL2.256` = 256-bit",
           "global accesses, `FMNMX3` = 3-input min/max, `CREDUX` = `redux.sync`, `STL`/`LDL` = local memory (spills).",
           "`multimem.st` has no SASS mnemonic of its own: it is an `STG.E.*.STRONG.SYS` to the multicast address",
           "(see `ptx_multimem_tma_excerpt.txt`). No tensor-core opcodes: the op is not a contraction.", "",
           "| kernel | instrs | LDG.256 | STG.256 | FMNMX3 | CREDUX | UBLKCP | SYNCS | LDGMC | STG.STRONG.SYS | LDG.STRONG.SYS | STL | LDL | F2I | I2F |",
           "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    flagship = None
    for f in funcs:
        name = f.split("\n")[0].strip()
        dem = run(["c++filt", name]).strip()
        ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M)
        c = collections.Counter()
        for i in ins:
            c["n"] += 1
            if i.startswith("LDG") and ".256" in i: c["ldg256"] += 1
            if i.startswith("STG") and ".256" in i: c["stg256"] += 1
            if i.startswith("FMNMX3"): c["fmnmx3"] += 1
            if i.startswith("CREDUX"): c["credux"] += 1
            if i.startswith("UBLKCP"): c["ublkcp"] += 1
            if i.startswith("SYNCS"): c["syncs"] += 1
            if i.startswith("LDGMC"): c["ldgmc"] += 1
            if i.startswith("STG") and "STRONG.SYS" in i: c["stgsys"] += 1
            if i.startswith("LDG") and "STRONG.SYS" in i: c["ldgsys"] += 1
            if i.startswith("STL"): c["stl"] += 1
            if i.startswith("LDL"): c["ldl"] += 1
            if i.startswith("F2I"): c["f2i"] += 1
            if i.startswith("I2F"): c["i2f"] += 1
        short = dem.replace("(anonymous namespace)::", "").replace("cgx::dev::", "").replace("cgx::", "")
        short = re.sub(r"\(.*", "", short).replace("void ", "")
        inv.append(f"| `{short}` | {c['n']} | {c['ldg256']} | {c['stg256']} | {c['fmnmx3']} | {c['credux']} | {c['ublkcp']} | {c['syncs']} | {c['ldgmc']} | {c['stgsys']} | {c['ldgsys']} | {c['stl']} | {c['ldl']} | {c['f2i']} | {c['i2f']} |")
        if "sra_kernel<float, 4, 2>" in dem or "sra_kernel<float, (int)4, (int)2>" in dem:
            flagship = f
    (OUT / "sass_inventory.md").write_text("\n".join(inv) + "\n")
    if flagship:
        (OUT / "sass_sra_kernel_f32_4bit_slice512.txt").write_text("Function : " + flagship)


def ptx():
    # the .so carries SASS only (code=sm_100a): regenerate the PTX of the fp32 instantiations
    src = ROOT / "torch_cgx_b200" / "csrc" / "kernels" / "sra_f32.cu"
    tmp = Path("/tmp/cgx_sra_f32.ptx")
    subprocess.run(["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--fmad=false", "-ptx",
                    "-o", str(tmp), str(src)], check=True, capture_output=True)
    keep, cur = [], None
    for line in tmp.read_text().splitlines():
        m = re.match(r"\.(?:visible |weak )*\.entry (\S+?)\(", line)
        if m:
            cur = run(["c++filt", m.group(1)]).strip()
        if cur and re.search(r"multimem\.|cp\.async\.bulk|mbarrier\.|prefetch\.global\.L2|ld\.global\.v8|st\.global\.v8|"
                             r"(min|max)\.NaN\.f32\s+%f\d+, %f\d+, %f\d+, %f\d+|redux\.sync|fence\.acq_rel\.sys|"
                             r"ld\.acquire\.sys|add\.rz\.f32", line):
            keep.append((re.sub(r"\(.*", "", cur).replace("cgx::dev::", ""), re.sub(r"%\w+", "%r", line.strip())))
    seen = collections.OrderedDict()
    for k, l in keep:
        if "<float, 4, 2>" in k or "<float, (int)4, (int)2>" in k:
            seen[(k, l)] = seen.get((k, l), 0) + 1
    out = ["PTX of the fp32 / 4-bit / 512-slice instantiations (nvcc -ptx csrc/kernels/sra_f32.cu, sm_100a):",
           "the lines that show the Blackwell data-movement features. Registers normalised to %r;",
           "count = occurrences inside that kernel.", ""]
    for (k, l), n in seen.items():
        out.append(f"{n:4d}  {k:45s} {l}")
    (OUT / "ptx_multimem_tma_excerpt.txt").write_text("\n".join(out) + "\n")


if __name__ == "__main__":
    g = ROOT / "gpurun_out"
    fused = g / "f_prof_fused_w1.ncu-rep"   # the newest capture wins
    if not fused.exists():
        fused = g / "a_prof_fused_w1.ncu-rep"
    if fused.exists():
        ncu_summary(fused, "ncu: fused SRA kernel, world = 1, 64 MiB fp32, 4-bit, bucket 512",
                    "World 1 runs only phase B (load -> min/max -> quantize -> pack -> self-decode -> store): 16 Mi elements, "
                    "64 MiB read + 64 MiB written (the writes mostly stay in the 126 MB L2, hence the small DRAM write figure).")
    if (g / "a_prof_quantize.ncu-rep").exists():
        ncu_summary(g / "a_prof_quantize.ncu-rep", "ncu: standalone quantize / dequantize item kernels, 64 MiB fp32, 4-bit, bucket 512",
                    "`quantize_items_kernel` reads 64 MiB and writes 8.25 MiB of wire; `dequantize_items_kernel` the reverse.")
    sass()
    ptx()
    print("profiles written to", OUT)
