"""profiles/<tag>_conv_layer_table.md + profiles/traffic.json from an ncu launch list of one bench step
(tools/gpu_profile_round.sh -> gpurun_out/step_launches_<tag>.csv).  Usage: python tools/make_profile_tables.py <tag>"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", "step_launches_%s.csv" % tag))) if len(r) > 14]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows[1:]:
    d = launches.setdefault(int(r[ix["ID"]]), {"kernel": r[ix["Kernel Name"]]})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3,
                                                              "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(r[ix["Metric Unit"]], 1)
names = ["stft_prep", "stft_gemm", "film"]
for k in range(7):
    names += ["enc%d.c1" % k, "enc%d.c2" % k]
for j in range(6):
    names += ["dec%d.up" % j, "dec%d.c1" % j, "dec%d.c2" % j]
names += ["mask_istft"]
# algorithmic GFLOP per clip of the 32 conv launches (2 * pixels * ncols * K)
enc_c = [(32, 32), (32, 64), (64, 128), (128, 256), (256, 384), (384, 384), (384, 384)]
dec_c = [(384, 384), (384, 384), (384, 256), (256, 128), (128, 64), (64, 32)]
H = [1024, 512, 256, 128, 64, 32, 32]
W = [512, 256, 128, 64, 32, 16, 8]
gflop = {}
for k, (ci, co) in enumerate(enc_c):
    px = H[k] * W[k]
    gflop["enc%d.c1" % k] = 2 * px * co * 9 * ci / 1e9
    gflop["enc%d.c2" % k] = 2 * px * co * (9 * co + (ci if k > 0 else 0)) / 1e9
for j, (ci, co) in enumerate(dec_c):
    lin, lo = 6 - j, 5 - j
    nup = 2 if j == 0 else 4
    gflop["dec%d.up" % j] = 2 * H[lin] * W[lin] * nup * co * ci / 1e9
    px = H[lo] * W[lo]
    gflop["dec%d.c1" % j] = 2 * px * co * 9 * 2 * co / 1e9
    gflop["dec%d.c2" % j] = 2 * px * co * (9 * co + 2 * co) / 1e9
B = 64
out = ["# Per-launch table of one bench step (64 clips x 10 s), ncu `gpu__time_duration.sum` + DRAM bytes (`%s_step_launches.csv`)" % tag, "",
       "| launch | kernel | us | DRAM read MB | DRAM write MB | GB/s | TFLOP/s (algorithmic) |", "|---|---|---|---|---|---|---|"]
tot_us = conv_bytes = all_bytes = 0.0
ordered = [d for _, d in sorted(launches.items())]
first = next(i for i, d in enumerate(ordered) if "stft_prep" in d["kernel"])
if first > 0:   # the capture window started mid-step: the front-end launches come from the following step
    ordered = ordered[first:first + 3] + ordered[:first]
ordered = ordered[:len(names)]
for d, nme in zip(ordered, names):
    us = d["gpu__time_duration.sum"]
    rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
    tot_us += us
    all_bytes += rd + wr
    kern = d["kernel"].split("(")[0].split("::")[-1].split("<")[0]
    if "conv_igemm_kernel" in d["kernel"]:
        kern = "conv_igemm_kernel<" + d["kernel"].split("conv_igemm_kernel<")[1].split(">")[0].replace("(int)", "") + ">"
    tf = ""
    if nme in gflop:
        conv_bytes += rd + wr
        tf = "%.0f" % (gflop[nme] * B / us * 1e3)
    out.append("| %s | `%s` | %.0f | %.0f | %.0f | %.0f | %s |" % (nme, kern, us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e3, tf))
out += ["", "Sum %.1f ms (serialised under ncu, cold caches).  Whole-step DRAM traffic %.1f GB = %.2f GB per clip (SURVEY.md §8d "
        "read-once/write-once minimum: 1.04 GB per clip)." % (tot_us / 1e3, all_bytes / 1e9, all_bytes / 1e9 / B)]
open(os.path.join(ROOT, "profiles", "%s_conv_layer_table.md" % tag), "w").write("\n".join(out) + "\n")
json.dump({"conv_unet_dram_bytes_per_step": conv_bytes, "all_kernels_dram_bytes_per_step": all_bytes, "batch": B,
           "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over the 32 conv launches of one step "
                     "(profiles/%s_step_launches.csv)" % tag},
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print("\n".join(out))
