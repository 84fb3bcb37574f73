"""Side-by-side per-launch times of gpurun_out/layer_times_<tag>.json files: python tools/cmp_layer_times.py tagA tagB ..."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tags = sys.argv[1:]
d = {t: json.load(open(os.path.join(ROOT, "gpurun_out", "layer_times_%s.json" % t))) for t in tags}
for k in d[tags[0]]["layers"]:
    base = d[tags[0]]["layers"][k]["ms"]
    print("%-9s " % k + "  ".join("%s %.4f (%+5.1f%%)" % (t, d[t]["layers"][k]["ms"], 100 * (d[t]["layers"][k]["ms"] / base - 1)) for t in tags)
          + "  %7.1f TF" % d[tags[0]]["layers"][k]["tflops"])
print("total    " + "  ".join("%s %.3f" % (t, d[t]["total_ms"]) for t in tags))
