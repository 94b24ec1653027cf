#!/usr/bin/env python
"""gpurun_out/emul_*.json (tests/tools/emulate_gpu.py on the B200 box) -> profiles/r02_precision_emulation.md"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
DESC = {
    "jitter": "fp32 + 6e-8 relative jitter on the conv inputs (noise floor between two fp32 runs of the SAME code)",
    "split": "product path: fp16 hi x hi + 2^-11 (hi x lo' + lo' x hi), 3 tensor-core products",
    "single": "one fp16 product",
    "fp8": "fp16 main + BOTH corrections in e4m3 (K-concatenated [Ah8|Al8'].[Wl8';Wh8]): 2 units of MMA work instead of 3",
    "fp8_e5a": "as fp8 with e5m2 activations",
    "hyb_a8": "Ah.[Wh|Wl'] in fp16 (N=128), only Al'.Wh in e4m3: 2.5 units",
    "hyb_w8": "Ah.Wh + Al'.Wh in fp16, only Ah.Wl' in e4m3: 2.5 units",
    "fp8_w2": "fp8 with two-term e4m3 weights (K x 2 on the corrections)",
    "fp8_a2": "fp8 with two-term e4m3 activations",
    "i8": "both corrections int8 (kind::i8): global activation step, per-output-channel weight step",
    "acts_only": "fp16 main + Al'.Wh only (weights rounded to fp16)",
    "weights_only": "fp16 main + Ah.Wl' only (activations rounded to fp16)",
    "wino_fp32": "Winograd F(2x2,3x3), fp32 transforms, fp32 products (transform rounding alone)",
    "wino_split": "Winograd F(2x2,3x3), fp32 transforms, fp16 x 3 split on the TRANSFORMED operands: 3/2.25 = 1.33 units",
}


def table(path, title, f):
    if not os.path.exists(path):
        return
    d = json.load(open(path))
    modes = [m for m in next(iter(d.values())) if m != "fp32_psnr"]
    f.write("## %s\n\nPer case: max over all iterate-map outputs of the relative L2 distance to the fp32 run of the same "
            "code (bar 1e-3), and the PSNR shift of the reconstruction (bar 0.05 dB).\n\n" % title)
    f.write("| scheme | " + " | ".join(d) + " |\n|---|" + "---|" * len(d) + "\n")
    f.write("| fp32 PSNR (dB) | " + " | ".join("%.3f" % d[c]["fp32_psnr"] for c in d) + " |\n")
    for m in modes:
        f.write("| `%s` | " % m + " | ".join("%.1e / %+.3f" % (d[c][m]["max_rel"], d[c][m]["dpsnr"]) for c in d) + " |\n")
    f.write("\n")
    return modes


with open(os.path.join(ROOT, "profiles", "r02_precision_emulation.md"), "w") as f:
    f.write("# Operand-format emulation on the B200 box (tests/tools/emulate_gpu.py; cuDNN fp32 convolutions of pre-rounded "
            "operands, TF32 off; full 256x256x8 solves, every iterate-map evaluation traced)\n\n")
    f.write("Cells: `max per-iterate rel-L2 / dPSNR (dB)` against the fp32 run.  Schemes apply to the 64->64 hidden "
            "layers only.\n\n")
    seen = []
    for path, title in (("emul_ffdnet.json", "DE-GAP-FFDnet (net_gray.pth stand-in), 180 iterations: 8 real measurements + 2 white-noise synthetic"),
                        ("emul_wino.json", "DE-GAP-FFDnet, Winograd F(2x2,3x3): 8 real measurements + 2 low-pass synthetic (the benchmarked data)"),
                        ("emul_cnn.json", "DE-GAP-CNN (cnn.ckpt), 100 iterations: 8 real measurements")):
        seen += table(os.path.join(G, path), title, f) or []
    f.write("## schemes\n\n")
    for m in dict.fromkeys(seen):
        f.write("* `%s`: %s\n" % (m, DESC.get(m, "")))
    log = os.path.join(G, "emul_synth.log")
    if os.path.exists(log):
        f.write("\n## noise floor of candidate synthetic generators (fp32 vs fp32 + jitter, and the product split)\n\n```\n")
        f.write(open(log).read())
        f.write("```\n")
print("wrote profiles/r02_precision_emulation.md")
