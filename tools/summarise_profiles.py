"""Cut the committed ncu summaries (profiles/) from the raw captures of tools/round_profile.sh in gpurun_out/.
usage: python tools/summarise_profiles.py r1c"""
import csv, json, os, shutil, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
WANT = ["Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summarise(rep, title, dst):
    hdr, units, rows = raw(rep)
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    lines = [title, f"report: gpurun_out/{os.path.basename(rep)} (scratch, not committed); ncu --set full --clock-control none --import-source on", ""]
    seen = set(); res = {}
    for r in rows:
        name = r[hdr.index("Kernel Name")].replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "").replace("b200::", "").split("(")[0]
        if name in seen:                                   # the same kernel launched again (PNG: once per strategy): keep every launch
            k = 2
            while f"{name} #{k}" in seen: k += 1
            name = f"{name} #{k}"
        seen.add(name)
        lines.append(f"== {name}")
        for w in WANT:
            if w in hdr: lines.append(f"   {w:78s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
        st = sorted(((float(r[hdr.index(s)]), s) for s in stall), reverse=True)[:6]
        lines.append("   top stall reasons (warps stalled per issue-active cycle): " + ", ".join(f"{s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, s in st))
        res[name] = {w: r[hdr.index(w)] for w in WANT if w in hdr}
        res[name]["units"] = {w: units[hdr.index(w)] for w in WANT if w in hdr}
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    return res


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def merged(*reps_titles_dst):
    res = {}
    for rep, title, dst in reps_titles_dst:
        if os.path.exists(rep): res.update(summarise(rep, title, dst))
    return res


t = summarise(os.path.join(G, f"{tag}_transform.ncu-rep"), f"{tag} -- transform kernels, one launch each inside `python bench.py --only-value` (one megabatch: 8 images of 3840x2160 per launch)", os.path.join(P, f"{tag}_ncu_transform_full.txt"))
e = summarise(os.path.join(G, f"{tag}_entropy.ncu-rep"), f"{tag} -- entropy kernels of one megabatch (8 images of 3840x2160) through b200_compress_batch (tools/profile_group.py 8)", os.path.join(P, f"{tag}_ncu_entropy_full.txt"))
o = merged((os.path.join(G, f"{tag}_png.ncu-rep"), f"{tag} -- PNG leg, one 4096x4096 RGBA image at --png-opt-level 3 (tools/profile_legs.py png): un-filter wavefront, K6 per strategy, K7 match / parse", os.path.join(P, f"{tag}_ncu_png_full.txt")),
           (os.path.join(G, f"{tag}i_png2.ncu-rep"), f"{tag} -- PNG leg, same image: K7 match after the bit-array rewrite, hash-chain candidates, DEFLATE coder kernels", os.path.join(P, f"{tag}_ncu_png2_full.txt")),
           (os.path.join(G, f"{tag}l_png3.ncu-rep"), f"{tag} -- PNG leg, same image, final K7 match (sixteen bytes per lane, one window array: 1.08 G warp instructions, 1.19 ms) and the register-array variant of the parse that was measured slower and dropped", os.path.join(P, f"{tag}_ncu_png3_full.txt")),
           (os.path.join(G, f"{tag}_webp.ncu-rep"), f"{tag} -- resize leg of 6000x4000 JPEG -> 1920-wide WebP (tools/profile_legs.py webp): YCbCr -> RGB, K3 Lanczos3 passes", os.path.join(P, f"{tag}_ncu_resize_full.txt")),
           (os.path.join(G, f"{tag}r_vp8tok.ncu-rep"), f"{tag} -- VP8 residual token pass on the device, 1920x1280 frame: per-macroblock masks, counting walk, writing walk (one thread per macroblock)", os.path.join(P, f"{tag}_ncu_vp8tok_full.txt")),
           (os.path.join(G, f"{tag}i_webp2.ncu-rep"), f"{tag} -- K8: RGB -> YUV and the VP8 wavefront kernel on a 1920x1280 frame", os.path.join(P, f"{tag}_ncu_vp8_full.txt")))
k = next(v for n, v in t.items() if "k_fused_same" in n)
fl = lambda key: float(k[key].replace(",", ""))
tr = {"source": f"profiles/{tag}_ncu_*_full.txt (ncu --set full --clock-control none, one launch per kernel; dram__bytes_read.sum + dram__bytes_write.sum)",
      "images_per_launch": {"jpeg kernels": 8, "png kernels": 1, "resize / vp8 kernels": 1}}
for name, v in list(t.items()) + list(e.items()) + list(o.items()):
    short = name.replace("void ", "").split("<")[0]
    if "dram__bytes_read.sum" in v:
        tr[short + "_bytes_per_launch"] = to_bytes(v["dram__bytes_read.sum"], v["units"]["dram__bytes_read.sum"]) + to_bytes(v["dram__bytes_write.sum"], v["units"]["dram__bytes_write.sum"])
tr["k_fused_same_pipes"] = {"warp_instructions_per_launch": fl("smsp__inst_executed.sum"), "issue_active_pct": fl("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                            "alu_pct": fl("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), "fma_pct": fl("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                            "fmaheavy_cycles_pct": fl("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"), "registers": fl("launch__registers_per_thread")}
json.dump(tr, open(os.path.join(P, "traffic.json"), "w"), indent=1)
for f in (f"{tag}_launches_bench.csv", f"{tag}_launches_group.csv", f"{tag}_launches_legs.csv", f"{tag}_bench.json", f"{tag}_bench_reference.json", f"{tag}_pytest_gpu.log",
          f"{tag}_coalesce0.json", f"{tag}_coalesce1.json"):
    if os.path.exists(os.path.join(G, f)): shutil.copy(os.path.join(G, f), os.path.join(P, f))
# per-kernel table of one megabatch
rows = list(csv.reader(open(os.path.join(G, f"{tag}_launches_group.csv"))))
i0 = next(i for i, r in enumerate(rows) if "Kernel Name" in r); h = rows[i0]
d = collections.OrderedDict()
for r in rows[i0 + 1:]:
    if len(r) <= h.index("Metric Value"): continue
    d.setdefault((int(r[h.index("ID")]), r[h.index("Kernel Name")].split("(")[0].replace("b200::", "")[:48]), {})[r[h.index("Metric Name")]] = float(r[h.index("Metric Value")].replace(",", ""))
L = list(d.items()); agg = collections.OrderedDict()
for (i, n), m in L[len(L) // 2:]:
    a = agg.setdefault(n, [0, 0.0, 0.0]); a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0); a[2] += m.get("smsp__inst_executed.sum", 0)
out = [f"{tag} -- second megabatch of tools/profile_group.py 8 under ncu (isolated, cold-cache durations; warp instructions executed)", ""]
for n, a in agg.items(): out.append(f"{n:50s} x{a[0]:3d} {a[1] / 1000:9.1f} us {a[2] / 1e6:9.1f} M warp-instr")
out.append(f"{'total':50s}      {sum(a[1] for a in agg.values()) / 1000:9.1f} us {sum(a[2] for a in agg.values()) / 1e6:9.1f} M warp-instr")
open(os.path.join(P, f"{tag}_group_kernels.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
