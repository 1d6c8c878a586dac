"""One-line summary of a bench.py output file (the JSON line may be preceded by library banners)."""
import json
import sys

for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        cfgd = d.get("config", {})
        print("%s B=%s N=%d: value %.1f  e2e %.1f  ms/step %.3f  frac %.3f  cpu %s  clocks %s" % (
            cfgd.get("net"), cfgd.get("images_per_step_per_gpu"), d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"],
            d.get("roofline", {}).get("frac", float("nan")), d.get("cpu_baseline", {}).get("value"), d.get("clocks")))
        break
else:
    sys.exit(1)
