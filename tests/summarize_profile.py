"""Turn an ncu report (gpurun_out/prof_<tag>.ncu-rep) + launch list into the committed text summaries under profiles/.
Runs in the build container (ncu -i needs no GPU)."""
import collections, csv, json, os, re, subprocess, sys

tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sm__cycles_elapsed.avg.per_second",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__inst_executed.avg.per_cycle_active", "l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum"]
vals = {a: (c, b) for a, b, c in zip(h, u, v)}
lines = [f"# ncu --set full, kernel fetch_kernel (step), report gpurun_out/prof_{tag}.ncu-rep", f"kernel: {vals.get('Kernel Name', ('?',''))[0]}"]
for k in keep:
    if k in vals:
        lines.append(f"{k:70s} {vals[k][0]:>16s} {vals[k][1]}")
def num(k):
    return float(vals[k][0].replace(",", ""))
unit = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}
dram = num("dram__bytes_read.sum") * unit[vals["dram__bytes_read.sum"][1]] + num("dram__bytes_write.sum") * unit[vals["dram__bytes_write.sum"][1]]
HAND = tag.startswith("hand")
ADROIT = tag.startswith("adroit")
KITCHEN = tag.startswith("kitchen")
ANT = tag.startswith("ant")
alg_b, alg_n = (1710, 2048) if HAND else ((1410, 2048) if ADROIT else ((1300, 2048) if KITCHEN else ((846, 1024) if ANT else (766, 4096))))
lines.append(f"dram bytes per launch (read+write): {dram:.0f}   [algorithmic: {alg_b} B x {alg_n} envs = {alg_b*alg_n}]")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
hdr = None; stalls = collections.Counter(); per = collections.Counter(); cur = None
def func_ranges(path):
    fr = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"^(HD|HDN|STAGE|static inline|__device__|__global__).*?\b(\w+)\s*\(", l)
        if m: fr.append((i, m.group(2)))
    return fr
cache = {}
for r in srows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) > 3 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        for a, b in zip(hdr, r):
            if a.startswith("stall_") and "(Not Issued)" not in a:
                try: stalls[a] += int(b)
                except ValueError: pass
        local = cur.replace("/root/repo", ROOT) if cur else None
        name = os.path.basename(cur or "?")
        if local and os.path.exists(local):
            fr = cache.setdefault(local, func_ranges(local))
            for s, n in fr:
                if s <= int(r[0]): name = os.path.basename(local) + ":" + n
        try: per[name] += int(r[hdr.index("Instructions Executed")])
        except ValueError: pass
tot = sum(stalls.values()) or 1
lines.append("\n# warp stall sampling (all samples), share of samples")
for k, n in stalls.most_common(10):
    lines.append(f"{k:28s} {100*n/tot:5.1f}%")
ti = sum(per.values()) or 1
lines.append(f"\n# executed warp instructions by source function (total {ti})")
for k, n in per.most_common(25):
    lines.append(f"{k:45s} {n:12d} {100*n/ti:5.1f}%")
open(os.path.join(out_dir, f"ncu_step_kernel_{tag}.txt"), "w").write("\n".join(lines) + "\n")
# the compute-side roofline figures bench.py attaches to its JSON line, per workload
def numk(k, default=None):
    try:
        return num(k)
    except Exception:
        return default
workload = {"hand": "hand_block_touch", "adroit": "adroit_hammer", "kitchen": "franka_kitchen", "ant": "antmaze_large"}.get(tag.split("_")[0], "fetch_pick_and_place")
tms = numk("gpu__time_duration.sum")
tunit = vals.get("gpu__time_duration.sum", ("", ""))[1]
tms = tms * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}.get(tunit, 1e-6) if tms is not None else None
fl = [numk("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"), numk("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"),
      numk("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum")]
roof = {"workload": workload, "ncu_source": f"profiles/ncu_step_kernel_{tag}.txt", "ncu_kernel_ms": tms, "envs_per_launch": alg_n,
        "dram_bytes_per_launch": dram, "algorithmic_bytes_per_launch": alg_b * alg_n,
        "issue_active_pct": numk("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": numk("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "fma_pipe_pct": numk("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "avg_active_lanes": numk("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "fp32_flop_per_launch": (fl[0] + fl[1] + 2 * fl[2]) if None not in fl else None,
        "local_load_store_inst": [numk("sass__inst_executed_local_loads"), numk("sass__inst_executed_local_stores")],
        "registers_per_thread": numk("launch__registers_per_thread"),
        "smem_per_block_bytes": (numk("launch__shared_mem_per_block_dynamic") or 0) * unit.get(vals.get("launch__shared_mem_per_block_dynamic", ("", "byte"))[1].split("/")[0], 1),
        "top_stalls": [[k.replace("stall_", ""), round(100 * n / tot, 1)] for k, n in stalls.most_common(5)]}
json.dump(roof, open(os.path.join(out_dir, f"roofline_{workload}.json"), "w"), indent=1)
json.dump({"dram_bytes_per_launch": dram, "source": f"profiles/ncu_step_kernel_{tag}.txt", "algorithmic_bytes_per_launch": alg_b * alg_n},
          open(os.path.join(out_dir, "traffic_hand.json" if HAND else ("traffic_adroit.json" if ADROIT else ("traffic_kitchen.json" if KITCHEN else ("traffic_ant.json" if ANT else "traffic.json")))), "w"))
# launch list
ll = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r: hh = r; st = i; break
    ki, vi = hh.index("Kernel Name"), hh.index("Metric Value")
    t = collections.Counter(); c = collections.Counter()
    for r in rows[st + 1:]:
        try: val = float(r[vi].replace(",", ""))
        except (ValueError, IndexError): continue
        t[r[ki][:90]] += val; c[r[ki][:90]] += 1
    s = sum(t.values())
    with open(os.path.join(out_dir, f"launches_{tag}.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 6 --warmup 3 (cold-cache, serialised: shares only)\n")
        for k, val in t.most_common(12):
            f.write(f"{k:92s} n={c[k]:4d} total_ns={val:14.0f} share={100*val/s:6.2f}%\n")
print("\n".join(lines[:30]))
