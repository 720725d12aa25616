"""Refresh profiles/traffic.json from an `ncu --set full` capture of the persistent ICP kernel on the default bench workload.
    ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/update_traffic.py /tmp/raw.csv <name> <passes>
The kernel-source hash stored with the figures is the one bench.py compares against HEAD (figures are used only on a match)."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {k: (u, float(v)) for k, u, v in zip(hdr, units, vals) if v.replace('.', '', 1).replace('e+', '', 1).replace('-', '', 1).isdigit()}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
def nbytes(key):
    u, v = d[key]
    return v * scale[u]
passes = float(sys.argv[3])
path = os.path.join(ROOT, "profiles", "traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
out = {
    "kernel_source_sha16": kernel_source_hash(),
    "icp_kernel_dram_bytes_per_launch": nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum"),
    "icp_kernel_lts_bytes_per_launch": d["lts__t_sectors.sum"][1] * 32.0,
    "icp_kernel_warp_instructions_per_hyp_iteration": d["smsp__inst_executed.sum"][1] / passes,
    "icp_kernel_issue_slots_busy_pct": d.get("sm__inst_issued.avg.pct_of_peak_sustained_active", d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", (None, None)))[1],
    "icp_kernel_hyp_iterations_in_capture": passes,
    "nn_query_kernel_dram_bytes_per_launch": old.get("nn_query_kernel_dram_bytes_per_launch"),
    "source": f"ncu --set full --clock-control none, profiles/{sys.argv[2]} (one icp_kernel launch of the default bench.py workload: 16 plots x 4096 hypotheses); "
              "warp-instructions = smsp__inst_executed.sum / hypothesis-iterations of the launch; L2 bytes = lts__t_sectors.sum x 32 B",
}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
