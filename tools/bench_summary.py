#!/usr/bin/env python
"""One-screen summary of bench.py JSON lines: `python tools/bench_summary.py file.json [...]`."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.load(open(path))
    except Exception as exc:
        print(path, "unreadable:", exc)
        continue
    if d.get("impl") == "reference":
        print(f"{path}: reference arm {d['value']:.2f} transients/s, {d['ms_per_step']:.1f} ms/step, {d['cpu_baseline']['cores']} threads, kind {d['cpu_baseline']['kind']}")
        continue
    r = d["roofline"]
    e = d.get("e2e") or {}
    print(f"{path}: N={d['n_gpus']} {d['value']:.0f} transients/s  {d['ms_per_step']*1e3:.1f} us/step  chain {r['chain_frac']:.3f} ({r['chain_frac_of_8tbs']:.3f} of 8 TB/s)"
          f"  fwd+bwd {d['fwd_bwd']['value']:.0f}/s ({d['fwd_bwd']['chain_frac']:.3f})")
    print("    stages: " + "  ".join(f"{s['kernel'][:12]} {s['ms']*1e3:.1f}us {s['frac']:.2f} [{s['bound'][:16]}]" for s in d["stages"]))
    ph = d["physical"]
    if ph.get("dram_gbs"):
        print(f"    physical: DRAM {ph['dram_gbs']:.0f} GB/s ({ph['dram_frac_of_peak']:.2f} of peak), fp32 {ph['fp32_tflops']:.1f} TFLOP/s")
    if e.get("value"):
        print(f"    e2e {e['value']:.0f}/s  link {e['link_gbs_per_direction_per_gpu']:.1f} GB/s per direction per GPU  e2e/link {e['frac_of_link']:.2f}")
    if d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        print(f"    cpu {c['value']:.2f}/s ({c['kind']}, {c['cores']} threads); same ops via torch on this GPU {c.get('same_ops_on_gpu_via_torch_transients_per_s')}")
    if d.get("model_path"):
        m = d["model_path"]
        print(f"    model path {m['ms_per_step']*1e3:.1f} us ({m['value']:.0f}/s); FeaturePropagation alone {m['feature_propagation_only_ms']*1e3:.1f} us")
    s3 = d.get("strong_cfg3")
    if s3 and "value" in s3:
        print(f"    strong cfg3: {s3['value']:.0f}/s  {s3['ms_per_step']:.3f} ms  per-GPU {s3['per_gpu']}  chain {s3['chain_frac']:.3f}  fwd+bwd {s3['fwd_bwd']['value']:.0f}/s")
    t4 = d.get("train_cfg4")
    if t4 and "ms_per_step" in t4:
        print(f"    train cfg4: {t4['ms_per_step']:.2f} ms/step (no allreduce {t4['ms_per_step_no_allreduce']:.2f}; allreduce alone {t4.get('allreduce_alone_ms')}; "
              f"overlap {t4.get('overlap')}); LCT part {t4['lct_part_ms']:.3f} ms")
    if d.get("latency_us", {}).get("eager"):
        print(f"    latency: eager {d['latency_us']['eager']:.0f} us, graph {d['latency_us']['cuda_graph']:.0f} us")
