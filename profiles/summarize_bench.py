"""Print the main numbers of a bench.py JSON line.    python profiles/summarize_bench.py file.json"""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print({k: d[k] for k in ("n_gpus", "value", "ms_per_step", "loss", "gpu_launches")}, d["clocks"])
e = d["e2e"]
print("e2e", round(e["value"]), "seq/s", round(e["ms_per_step"], 4), "ms | encoder output on device:",
      round(e["encoder_output_on_device"]["value"]), "seq/s")
r = d["roofline"]
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k not in ("peak_source", "kernel")})
print("compute_losses ms", round(d["compute_losses"]["ms_per_step"], 4), "| module_api ms", round(d["module_api"]["ms_per_step"], 4))
if d.get("dp_collectives"):
    print("dp_collectives", d["dp_collectives"])
if d.get("per_rank_region_ms"):
    print("per_rank", d["per_rank_region_ms"])
for name, leg in (d.get("configs") or {}).items():
    if not leg:
        continue
    if name == "cfg3_ccl_sampled":
        print(name, round(leg["value"]), leg["unit"], round(leg["ms_per_step"], 4), "ms; roofline frac", round(leg["roofline"]["frac"], 3),
              "kernel_ms", round(leg["roofline"]["kernel_ms"], 4))
    elif name == "cfg5_points":
        for p in leg["points"]:
            print(name, p["loss"], p["queries"], "x", p["candidates"], "fused", round(p["fused_ms"], 3), "ms (kernel", round(p["fused_kernel_ms"], 3),
                  ") materialised", round(p["materialised_logits_ms"], 2), "ms speedup", round(p["speedup"], 1), "frac burst",
                  round(p["roofline"]["frac"], 3), "sustained", round(p["roofline"]["frac_of_sustained_peak"], 3))
    elif name == "encoder_train_step":
        print(name, round(leg["value"]), leg["unit"], round(leg["ms_per_step"], 4), "ms; encoder forward (eager)", round(leg["encoder_forward_eager_ms"], 4), "ms")
    else:
        for kname, kk in leg.items():
            print(name, kname, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in kk["roofline"].items() if a not in ("note",)})
rt = d.get("retrieval")
if rt:
    print("retrieval", round(rt["value"]), rt["unit"], rt.get("exchange"), "sharded_equals_unsharded:", rt.get("sharded_equals_unsharded"))
    for p in rt["points"]:
        print("  U", p["queries"], "ms", round(p["ms_per_batch"], 4), "q/s", round(p["queries_per_s"]), "min/max", [round(x, 3) for x in p["ms_per_batch_min_max"]],
              p["roofline"]["bound"], "frac", round(p["roofline"]["frac"], 3), "kernel_ms", round(p["roofline"]["kernel_ms"], 4))
    if "cpu_baseline" in rt:
        print("  cpu", rt["cpu_baseline"]["value"], rt["cpu_baseline"]["unit"])
print("cpu_baseline", d.get("cpu_baseline"))
