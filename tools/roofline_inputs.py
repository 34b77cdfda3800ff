"""Extracts the roofline inputs of one kernel from an .ncu-rep (ncu --set full) into profiles/roofline_inputs.json:
    python tools/roofline_inputs.py <x.ncu-rep> <key> <units_per_launch> [summary.txt] [src,files]
key = "<kernel>/<config>", e.g. "sample_seatfirst_kernel<true>/c2"; units = raw deals (or records) the captured launch
processed.  Records W_int = sass__thread_inst_executed_true / units, DRAM bytes, issue utilisation and the hash of the
library sources the capture belongs to (bench.py flags a mismatch as stale)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import SAMPLER_SOURCES, source_hash  # noqa: E402

rep, key, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
d = dict(zip(hdr, rows[2]))


def num(name):
    return float(d[name].replace(",", ""))


def bytes_of(name):
    unit = rows[1][hdr.index(name)]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return num(name) * scale


def time_ms(name):
    unit = rows[1][hdr.index(name)]
    return num(name) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]


ent = {
    "kernel": d["Kernel Name"],
    "w_int_per_deal": num("sass__thread_inst_executed_true_per_opcode") / units if "sass__thread_inst_executed_true_per_opcode" in d
    else num("smsp__thread_inst_executed.sum") / units,
    "deals_per_launch": units,
    "dram_bytes_per_launch": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
    "launch_ms_under_ncu": time_ms("gpu__time_duration.sum"),
    "issue_active_pct": num("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    "registers": num("launch__registers_per_thread"),
    "source": sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep),
}
files = tuple(sys.argv[5].split(",")) if len(sys.argv) > 5 else SAMPLER_SOURCES
ent["src_files"] = list(files)
ent["src_hash"] = source_hash(files)
for extra in ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active"):
    if extra in d:
        ent[extra.split(".")[0]] = num(extra)
path = os.path.join(ROOT, "profiles", "roofline_inputs.json")
table = json.load(open(path)) if os.path.exists(path) else {}
table[key] = ent
json.dump(table, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(ent))
