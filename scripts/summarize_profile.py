"""Turn ncu outputs (gpurun_out/) into the small text/JSON summaries committed under profiles/.

    python scripts/summarize_profile.py launches gpurun_out/launches.csv profiles/r1_launches_c2m.txt
    python scripts/summarize_profile.py full gpurun_out/prof.ncu-rep profiles/r1_ncu_full_c2m.txt [profiles/kernel_traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys

NAMES = {"prologue_w": "mh_prologue_w", "norm_backward_w": "mh_norm_backward_w"}
# tc_kernel<MODE, VARIANT>: MODE 0 FWD, 1 FWDS (forward + stash), 2 BWD_G, 3 DX, 4 DW (see tc_head.cu)
TC_MODES = ["mh_tc_forward", "mh_tc_forward", "mh_tc_backward_g", "mh_tc_backward_dx", "mh_tc_backward_dw_fused"]


def api_name(kernel):
    import re
    if "tc_kernel_dxdw" in kernel:
        return "mh_tc_backward_dxdw"
    m = re.search(r"tc_kernel<\(?(?:int\))?(\d)", kernel)
    if m:
        return TC_MODES[int(m.group(1))]
    for k, v in NAMES.items():
        if k in kernel:
            return v
    return None


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        a = agg.setdefault(r[kn][:110], [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (one fused ArcFace head step, B=1024, C=2,000,000)\n")
        f.write(f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes. source: {src}\n")
        f.write(f"{'us':>10s} {'n':>3s} {'share':>7s}  kernel\n")
        for k, (v, n) in agg.items():
            f.write(f"{v:10.1f} {n:3d} {100 * v / tot:6.1f}%  {k}\n")
        f.write(f"{tot:10.1f}     total\n")
    print(open(dst).read())


WANT = [
    "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__cycles_elapsed.max",
]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def full(src, dst, traffic_json=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    kn = h.index("Kernel Name")
    traffic = {}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on (B=1024, C=2,000,000, ArcFace); source: {src}\n")
        for r in rows[2:]:
            f.write(f"\n== {r[kn][:120]}\n")
            for w in WANT:
                if w in h:
                    i = h.index(w)
                    f.write(f"   {w:72s} {r[i]:>16s} {units[i]}\n")
            ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            t = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            f.write(f"   {'dram traffic per launch (read+write)':72s} {t / 1e9:16.3f} GB\n")
            n = api_name(r[kn])
            if n:
                traffic[n] = t
    print(open(dst).read())
    if traffic_json:
        json.dump(traffic, open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
