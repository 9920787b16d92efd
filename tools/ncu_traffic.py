#!/usr/bin/env python
"""DRAM traffic per launch of the encoder GEMMs from an `ncu --set full` report -> profiles/r02_gemm_traffic.json
(bench.py copies `dram_bytes_per_launch` into roofline.traffic and `source` into roofline.traffic_source).
    python tools/ncu_traffic.py gpurun_out/r02_encoder_full.ncu-rep "<commit / command the capture belongs to>" """
import csv, io, json, os, subprocess, sys


def to_bytes(v, u):
    f = float(v.replace(",", ""))
    return f * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)


def main(path, note):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    per = []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if "gemm_tc2_kernel" not in name:
            continue
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        t, tu = float(r[idx["gpu__time_duration.sum"]].replace(",", "")), units[idx["gpu__time_duration.sum"]]
        us = t * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(tu, 1.0)
        per.append({"kernel": name.replace("void ", "").replace("ac::", "").split("(CUtensorMap")[0][:80], "us": round(us, 1),
                    "dram_read": rd, "dram_write": wr})
    # one layer = QKV, out-proj, FFN1, FFN2: keep whole groups of four consecutive launches
    per = per[:len(per) // 4 * 4]
    mean = sum(p["dram_read"] + p["dram_write"] for p in per) / max(len(per), 1)
    res = {"dram_bytes_per_launch": mean,
           "source": f"ncu --set full --clock-control none, {note}: {len(per)} consecutive encoder GEMM launches (QKV, out-proj, FFN1, FFN2 per layer), "
                     "B=512 S=128; mean of dram__bytes_read.sum + dram__bytes_write.sum per launch",
           "algorithmic_bytes_per_launch_note": "operands in + results out at B*S = 65536: QKV 101 + 302 = 403 MB, out-proj 101 + 201 (fp32 sums read) + 201 (written) + 101 (fp16 copy) = 604 MB, FFN1 101 + 403 = 503 MB, FFN2 403 + 201 + 201 + 101 = 906 MB; mean 604 MB (part of the written bytes is still in the 126 MB L2 when a kernel ends)",
           "per_launch": per}
    json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_gemm_traffic.json"), "w"), indent=1)
    print(json.dumps({k: v for k, v in res.items() if k != "per_launch"}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
