"""profiles/ncu_facts.json from the text summaries of the `ncu --set full` captures under profiles/ (dev tool)."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
facts = {"source": "ncu --set full --clock-control none, one launch inside an eager step at batch 16 (profiles/r02_ncu_full_*.txt)"}
for key, name in (("fc2_gemm", "fc2"), ("fc1_gemm", "fc1"), ("qkv_gemm", "qkv"), ("out_proj_gemm", "out_proj"), ("flash_attn", "flash")):
    txt = open(os.path.join(ROOT, "profiles", f"r02_ncu_full_{name}.txt")).read()
    def val(metric):
        m = re.search(re.escape(metric) + r": ([\d.,]+) (\S*)", txt)
        v = float(m.group(1).replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(2), 1)
    facts[key + "_dram_bytes"] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
    facts[key + "_tensor_pipe_pct"] = val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    facts[key + "_ncu_us"] = val("gpu__time_duration.sum")
json.dump(facts, open(os.path.join(ROOT, "profiles", "ncu_facts.json"), "w"), indent=1)
print(json.dumps(facts, indent=1))
