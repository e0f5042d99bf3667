"""Summarise an `ncu --set full` report of one ESPCN training step (k_conv_rs | k_conv_sl | k_tc_wgrad instances, in launch order)
into profiles/: a JSON keyed by "<C-ABI call>|<layer>" (bench.py reads `traffic` from it) and a markdown table.

    python tools/ncu_summary.py <report.ncu-rep> <tag>
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
GENERIC = len(sys.argv) > 3  # third argument = workload name: no layer mapping, table only (profiles/<tag>_ncu_full_<workload>.md)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}

# launch order of the tensor-core kernels inside one ESPCN cfg2 step (fwd L1..L3, then bwd L3..L1)
L1 = "N128 Cin3 64x64 Cout64 k5 s1 p0 ps1"
L2 = "N128 Cin64 60x60 Cout32 k3 s1 p0 ps1"
L3f = "N128 Cin32 58x58 Cout3 k3 s1 p0 ps4"
L3b = "N128 Cin32 58x58 Cout48 k3 s1 p0 ps1"
ORDER = [("srb_conv_fprop", L1), ("srb_conv_fprop", L2), ("srb_conv_fprop", L3f), ("srb_conv_wgrad", L3b),
         ("srb_conv_dgrad", L3b), ("srb_conv_wgrad", L2), ("srb_conv_dgrad", L2), ("srb_conv_wgrad", L1)]


def f(r, k, d=0.0):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return d


def unit(k):
    return rows[1][ix[k]] if k in ix else ""


out, md = {}, ["| call | layer | kernel | grid | time us | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe % | mem->tensor % | L1/TEX % | regs | dyn smem KB |",
               "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for pos, r in enumerate(data):
    if not GENERIC and pos >= len(ORDER):
        break
    call, layer = (("-", "launch %d" % pos) if GENERIC else ORDER[pos])
    name = r[ix["Kernel Name"]]
    kern = next((k for k in ("k_conv_rs", "k_conv_sl", "k_tc_wgrad", "k_wgrad_finish") if k in name), name[:30])
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = f(r, "dram__bytes_read.sum") * scale.get(unit("dram__bytes_read.sum"), 1.0)
    wr = f(r, "dram__bytes_write.sum") * scale.get(unit("dram__bytes_write.sum"), 1.0)
    tus = f(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit("gpu__time_duration.sum"), 1.0)
    tens = f(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
    memt = f(r, "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    l1 = f(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed")
    rec = {"kernel": kern, "grid": r[ix["Grid Size"]], "time_us": tus, "dram_read_bytes": rd, "dram_write_bytes": wr,
           "traffic": rd + wr, "dram_gbs": (rd + wr) / tus * 1e-3 if tus else None, "tensor_pipe_pct": tens,
           "mem_tensor_pct": memt, "l1tex_pct": l1, "regs": f(r, "launch__registers_per_thread"),
           "dyn_smem_kb": f(r, "launch__shared_mem_per_block_dynamic")}
    out[call + "|" + layer] = rec
    md.append("| %s | %s | %s | %s | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %d | %.0f |" % (
        call, layer, kern, rec["grid"], tus, rd / 1e6, wr / 1e6, rec["dram_gbs"] or 0, tens, memt, l1, rec["regs"], rec["dyn_smem_kb"]))
if GENERIC:
    open(os.path.join(ROOT, "profiles", "%s_ncu_full_%s.md" % (tag, sys.argv[3])), "w").write(
        "# ncu --set full, %s: consecutive tensor-core kernel launches of one training step\n\n"
        "Captured with `--clock-control none`, kernels replayed and serialised (cold caches).\n\n" % sys.argv[3] + "\n".join(md) + "\n")
    print("\n".join(md))
    sys.exit(0)
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
open(os.path.join(ROOT, "profiles", "%s_ncu_full_espcn.md" % tag), "w").write(
    "# ncu --set full, one ESPCN cfg2 training step (tensor-core kernels in launch order)\n\n"
    "Captured with `--clock-control none`, kernels replayed and serialised (cold caches): compare SHARES, not absolutes.\n\n" + "\n".join(md) + "\n")
print("\n".join(md))
