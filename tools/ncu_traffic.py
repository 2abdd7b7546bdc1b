"""profiles/r2_ncu_traffic.json — DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels
bench.py reports a roofline for, taken from `ncu --set full` captures (gpurun_out/*.ncu-rep; run where ncu is installed).
bench.py copies these figures into the `traffic` fields of its JSON line.

    python tools/ncu_traffic.py <tag>        # reads gpurun_out/<tag>_ncu_pxn_layer1.ncu-rep, <tag>_ncu_fused.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def to_bytes(r, key):
        v, u = float(r[col[key]].replace(",", "")), units[col[key]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    res = []
    for r in rows[2:]:
        res.append({"kernel": r[col["Kernel Name"]].split("(")[0],
                    "us": float(r[col["gpu__time_duration.sum"]].replace(",", "")) *
                          {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(units[col["gpu__time_duration.sum"]].lower(), 1.0),
                    "read": to_bytes(r, "dram__bytes_read.sum"), "write": to_bytes(r, "dram__bytes_write.sum"),
                    "tensor_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]])})
    return res


def main():
    tag = sys.argv[1]
    out = {"source": f"ncu --set full --clock-control none, gpurun_out/{tag}_ncu_*.ncu-rep (tools/gpu_visit.sh), "
                     "per launch: dram__bytes_read.sum + dram__bytes_write.sum", "launches": {}}
    pxn = launches(os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_pxn_layer1.ncu-rep"))
    conv = [l for l in pxn if l["kernel"].endswith("pxn_kernel")]
    if conv:
        out["tapgemm_total_bytes_per_launch"] = sum(l["read"] + l["write"] for l in conv) / len(conv)
        out["tapgemm_geometry"] = ("pxn_kernel at the layer1 geometry (64->64, 3x3, 32x32 maps, batch 1026): fprop "
                                   "(x 134.5 MB read, z fp32 269 MB written) and dgrad (dz read, dx bf16 134.5 MB written); "
                                   "writes still resident in the 126 MB L2 at kernel end are not counted by the DRAM "
                                   "counters")
    hbm = {}
    for l in pxn:
        for name in ("affine_apply_kernel", "column_reduce_kernel", "bwd_dz_kernel"):
            if name in l["kernel"]:
                hbm[name] = {"dram_bytes": l["read"] + l["write"], "us": l["us"]}
    if hbm:
        out["hbm_passes"] = hbm
    fused = launches(os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_fused.ncu-rep"))
    if fused:
        out["passport_fused_bytes_per_launch"] = fused[0]["read"] + fused[0]["write"]
        out["passport_fused_tensor_pipe_pct"] = fused[0]["tensor_pct"]
    seq = os.path.join(ROOT, "gpurun_out", "r2e_ncu_sequence.ncu-rep")
    if os.path.exists(seq):
        tg = [l for l in launches(seq) if "tapgemm_kernel" in l["kernel"]]
        if tg:
            out["passport_layer_bytes_per_launch"] = tg[0]["read"] + tg[0]["write"]
    for l in pxn + fused:
        out["launches"].setdefault(l["kernel"], []).append({k: l[k] for k in ("us", "read", "write", "tensor_pct")})
    with open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
