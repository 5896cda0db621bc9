"""Copy the round's evidence from gpurun_out/ (scratch) into profiles/ (tracked): python tools/collect_profiles.py r01m"""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
FRAMES = int(sys.argv[2]) if len(sys.argv) > 2 else 1000          # frames of the `ncu --set full` capture
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for w in ["1080p", "4k", "480p", "4k-q1", "enc-1080p", "reference", "enc_reference", "1080p_2gpu", "1080p_8gpu", "1080p-8192_8gpu"]:
    src = os.path.join(G, f"bench_{w}.log")
    if os.path.exists(src):
        lines = [l for l in open(src).read().splitlines() if l.startswith("{")]
        if lines:
            with open(os.path.join(P, f"{tag}_bench_{w}.json"), "w") as f:
                f.write(lines[-1] + "\n")
            print("bench", w, json.loads(lines[-1])["value"])
src = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, f"{tag}_ncu_launches_bench_1080p_steps2.csv"))
rep = os.path.join(G, f"prof_{tag}.ncu-rep")
if os.path.exists(rep):
    out = os.path.join(P, f"{tag}_ncu_full_pipeline_1080p_{FRAMES}f.csv")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, out], check=True, stdout=subprocess.DEVNULL)
    rows = list(csv.reader(open(out)))
    import re
    names = [re.sub(r"<.*", "", n.replace("void ", "")).strip() for n in rows[0][2:]]      # "void k_x<1>" -> "k_x"
    get = lambda m: [float(x.replace(",", "")) for x in next(r for r in rows if r[0] == m)[2:]]
    unit = lambda m: next(r for r in rows if r[0] == m)[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = [v * scale[unit("dram__bytes_read.sum")] for v in get("dram__bytes_read.sum")]
    wr = [v * scale[unit("dram__bytes_write.sum")] for v in get("dram__bytes_write.sum")]
    traffic = {"1080p": {n: {"dram_bytes_per_frame": (a + b) / FRAMES, "read": a / FRAMES, "write": b / FRAMES,
                             "source": f"profiles/{os.path.basename(out)} (tools/profile_run.py --frames {FRAMES})"}
                         for n, a, b in zip(names, rd, wr)}}
    json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
    print("traffic", {n: round(v["dram_bytes_per_frame"]) for n, v in traffic["1080p"].items()})
