"""profiles/dominant_kernel_traffic.json from in-model `ncu --set full` captures (run on the CPU box after round_records.sh):

    python tools/make_traffic_json.py <tag>      # reads gpurun_out/<tag>_c128_k11_block.ncu-rep (+ _narrow_conv2, _wn_layer)

dram bytes = dram__bytes_read.sum + dram__bytes_write.sum per launch; the dominant family's six launches are listed one by
one and averaged (bench.py puts the average into roofline.traffic)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
B, C, L = 16, 128, 65536
E = 4.0 * B * C * L  # bytes of one fp32 tensor / one two-plane operand image of the C=128 stage


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for data in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, data):
            d[h] = (u, v)
        res.append(d)
    return res


def num(d, key):
    u, v = d[key]
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "%": 1.0}.get(u, 1.0)
    return x * scale


def launch(d):
    return {"dram_read_bytes": num(d, "dram__bytes_read.sum"), "dram_write_bytes": num(d, "dram__bytes_write.sum"),
            "duration_us_under_ncu": num(d, "gpu__time_duration.sum"),
            "tensor_pipe_active_pct": num(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")}


fam = {}
blk = raw(os.path.join(ROOT, "gpurun_out", f"{tag}_c128_k11_block.ncu-rep"))
W = 4.0 * C * C * 11
# the k=11 block is the LAST block of its stage: conv1 image -> image; conv2 + image residual -> image; the last conv2 also
# reads the fp32 running sum and writes only the next upsampler's image
what = [("conv1 d1: image -> image", 2 * E + W), ("conv2: + image residual -> image", 3 * E + W),
        ("conv1 d3", 2 * E + W), ("conv2: + image residual -> image", 3 * E + W), ("conv1 d5", 2 * E + W),
        ("conv2: + image residual + fp32 running sum -> image of the stage output / 3 (last block of the stage)", 4 * E + W)]
ls = []
for d, (w, alg) in zip(blk, what):
    e = {"what": w}
    e.update(launch(d))
    e["algorithmic_bytes"] = alg
    ls.append(e)
fam["tc:cin128:k11"] = {
    "dram_bytes_per_launch": sum(x["dram_read_bytes"] + x["dram_write_bytes"] for x in ls) / len(ls),
    "algorithmic_bytes_per_launch_average": sum(x["algorithmic_bytes"] for x in ls) / len(ls), "launches": ls,
    "what": "the six launches of ResBlock k=11 of the C=128 stage (the dominant family), captured IN THE MODEL at 16x80x1024 "
            f"(tools/round_records.sh {tag}: ncu --set full -k regex:conv_tc_kernel -s <first> -c 6 python tools/ncu_target.py); "
            "dram bytes = read + write, averaged over the six"}
p = os.path.join(ROOT, "gpurun_out", f"{tag}_narrow_conv2.ncu-rep")
if os.path.exists(p):
    d = raw(p)[0]
    e = launch(d)
    e["dram_bytes_per_launch"] = e["dram_read_bytes"] + e["dram_write_bytes"]
    e["achieved_gbs"] = e["dram_bytes_per_launch"] / (e["duration_us_under_ncu"] * 1e-6) / 1e9
    e["what"] = "resblock conv2, C=64 k=7 (image-only residual stream), in the model: xt image + residual image -> image"
    fam["tc:cin64:k7"] = e
p = os.path.join(ROOT, "gpurun_out", f"{tag}_wn_layer.ncu-rep")
if os.path.exists(p):
    d = raw(p)[0]
    e = launch(d)
    e["dram_bytes_per_launch"] = e["dram_read_bytes"] + e["dram_write_bytes"]
    e["what"] = "the 16-layer encoder WN stack in one launch (wn_layer_kernel, n_layers = 16), in the model"
    fam["tc:wn_layer"] = e
out = {"source": f"ncu --set full --clock-control none --import-source on, launches captured inside SynthesizerTrn.infer at 16x80x1024 "
                 f"(tools/round_records.sh {tag}; summaries profiles/{tag}_ncu_*.txt); per launch", "families": fam}
json.dump(out, open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json"), "w"), indent=1)
print(json.dumps({k: (v["dram_bytes_per_launch"], v.get("algorithmic_bytes_per_launch_average")) for k, v in fam.items()}))
