"""profiles/ncu_facts.json from ncu --set full captures (dev tool): ncu_facts.py gemm.ncu-rep flash.ncu-rep"""
import csv, json, os, subprocess, sys
def rows(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    return [dict(zip(r[0], x)) for x in r[2:]], dict(zip(r[0], r[1]))
def nbytes(d, u, k):
    v = float(d[k].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
facts = {"source": "ncu --set full --clock-control none, batch 16 (profiles/r01_ncu_full_*.txt)"}
g, gu = rows(sys.argv[1])
d = [x for x in g if "EpiF16<(int)1>" in x["Kernel Name"] or "EpiF16<1>" in x["Kernel Name"]][0]
facts["fc1_gemm_dram_bytes"] = nbytes(d, gu, "dram__bytes_read.sum") + nbytes(d, gu, "dram__bytes_write.sum")
facts["fc1_gemm_tensor_pipe_pct"] = float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])
f, fu = rows(sys.argv[2])
d = [x for x in f if "flash_attn" in x["Kernel Name"]][0]
facts["flash_attn_dram_bytes"] = nbytes(d, fu, "dram__bytes_read.sum") + nbytes(d, fu, "dram__bytes_write.sum")
facts["flash_attn_tensor_pipe_pct"] = float(d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"])
json.dump(facts, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_facts.json"), "w"), indent=1)
print(facts)
