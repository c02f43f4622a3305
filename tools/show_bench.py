import json, sys
for f in sys.argv[1:]:
    d = json.load(open(f))
    if "roofline" not in d:
        print(f, round(d["value"], 1), d["unit"], d.get("cpu_baseline"))
        continue
    r = d["roofline"]
    print(f, round(d["value"]), "Mrays/s", round(d["sample_bounces_per_s"] / 1e9, 3), "Gsb/s", round(d["samples_per_s"] / 1e6), "Msamples/s e2e",
          round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 1), r["kernel"], "frac", round(r["frac"], 3),
          {k: round(v, 1) for k, v in r["stage_ms_per_step"].items()}, "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 1), "launches", d["gpu_launches"], d["clocks"])
