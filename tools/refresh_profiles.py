"""Copies the outputs of tools/round_profile.sh <tag> from gpurun_out/ into profiles/ under the round's names:
    python tools/refresh_profiles.py r02f r02
bench line, launch lists, step timeline, kernel tables (<round>_kernels.md) and the DRAM traffic per matvec launch
(<round>_matvec_traffic.json, read by bench.py for `roofline.traffic`)."""
import json
import os
import shutil
import subprocess
import sys

tag, rnd = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
line = open(os.path.join(src, f"{tag}_bench.json")).read().strip().splitlines()[-1]
json.dump(json.loads(line), open(os.path.join(dst, f"{rnd}_bench_line.json"), "w"), indent=1)
for wl in ("h2s", "h2o", "ocs_batch"):
    shutil.copy(os.path.join(src, f"{tag}_launches_{wl}.csv"), os.path.join(dst, f"{rnd}_launches_{wl}.csv"))
if os.path.exists(os.path.join(src, f"{tag}_step_timeline.txt")):
    with open(os.path.join(dst, f"{rnd}_step_timeline.txt"), "w") as f:
        f.write("# Device timelines of the device-resident bench step (tools/gpu_timeline.py: CUPTI through torch.profiler, 4 steps\n"
                "# after 8 warm-up steps; no profiler-timed number is a bench value -- the 'unprofiled' line is the same loop timed\n"
                "# with CUDA events first)\n")
        f.write(open(os.path.join(src, f"{tag}_step_timeline.txt")).read())


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    import csv
    rows = list(csv.reader(out.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    return {k: (v, u) for k, u, v in zip(h, units, vals)}


def num(m, key):
    v, u = m[key]
    v = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v * scale


traffic, md = {}, []
for wl, kern, rep in (("h2s", "k_matvec_dmma", "dmma"), ("h2o", "k_matvec_tiled", "tiled"), ("ocs_batch", "k_matvec_lin", "lin")):
    path = os.path.join(src, f"{tag}_{rep}.ncu-rep")
    m = metrics(path)
    traffic[wl] = {"kernel": kern, "dram_bytes_read": num(m, "dram__bytes_read.sum"), "dram_bytes_write": num(m, "dram__bytes_write.sum"),
                   "duration_us": num(m, "gpu__time_duration.sum"),
                   "capture": f"gpurun_out/{tag}_{rep}.ncu-rep (ncu --set full --clock-control none, one full-batch launch, "
                              "tools/round_profile.sh)"}
    md.append(f"## {rep} (ncu --set full --clock-control none, one full-batch launch)\n" +
              subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), "kernel", path], capture_output=True,
                             text=True).stdout)
path = os.path.join(src, f"{tag}_recur.ncu-rep")
md.append("## recur (k_recur_gram inside the H2S bench step)\n" +
          subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), "kernel", path], capture_output=True,
                         text=True).stdout)
json.dump(traffic, open(os.path.join(dst, f"{rnd}_matvec_traffic.json"), "w"), indent=1)
open(os.path.join(dst, f"{rnd}_kernels.md"), "w").write("\n".join(md))
for wl in ("h2s", "h2o", "ocs_batch"):
    print(f"### {wl}")
    print(subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_summary.py"), "launches",
                          os.path.join(dst, f"{rnd}_launches_{wl}.csv")], capture_output=True, text=True).stdout)
