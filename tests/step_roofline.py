"""Whole-step roofline table: per kernel class of one 80-frame try-on step, ncu duration (launch list) vs the
algorithmic bytes / FLOPs of that class.  usage: step_roofline.py launches.csv passes  ->  markdown on stdout."""
import csv
import json
import os
import sys
from collections import defaultdict

F, H, W = 80, 256, 192
PX = H * W
HBM, TF = 6584.8, 1457.6
try:
    pk = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    HBM, TF = pk["hbm_gbs"], pk["bf16_tflops_sustained"]
except Exception:  # noqa: BLE001
    pass


def planes(n_elems):  # hi + lo 16-bit planes
    return 4 * n_elems


# U-Net feature maps (channels, pixels per frame) after each down conv / each up conv
ENC = [(64, PX // 4), (128, PX // 16), (256, PX // 64), (512, PX // 256), (512, PX // 1024), (512, PX // 4096)]
DEC = [(512, PX // 1024), (512, PX // 256), (256, PX // 64), (128, PX // 16), (64, PX // 4), (4, PX)]
norm_elems = F * (sum(c * p for c, p in ENC) + sum(c * p for c, p in DEC))
stats_elems = F * (sum(c * p for c, p in ENC[1:5]) + sum(c * p for c, p in DEC))
# decoder tap-stacked partials: 9*Cout floats per LOW-res pixel in, 4*Cout out
gather_bytes = F * sum(4 * (9 * c * (p // 4) + c * p) for c, p in DEC)
ALG = {  # kernel-name fragment -> (algorithmic MB per step, what)
    "nchw_s2d_planes": ((F * 22 * PX * 4 + planes(F * 129 * 97 * 128) + F * 10 * PX * 4 + planes(F * 129 * 97 * 64)) / 1e6,
                        "f32 NCHW in + s2d planes out (GMM person 22 ch, U-Net person+cloth 10 ch)"),
    "nchw_im2col_planes": ((F * 3 * PX * 4 + planes(F * (PX // 4) * 64)) / 1e6, "cloth 3 ch in + im2col planes (K 48 -> 64) out"),
    "instnorm_apply": ((4 * norm_elems + planes(norm_elems)) / 1e6, "f32 conv output in + normalised / activated planes out"),
    "instnorm_stats": (4 * stats_elems / 1e6, "f32 conv output in"),
    "upconv3x3_gather": (gather_bytes / 1e6, "9*Cout partials per low-res pixel in + f32 output"),
    "tps_grid_sample": (F * 6 * PX * 4 / 1e6, "cloth in + warped cloth out (grid never materialised)"),
    "tom_compose": (F * (4 + 3 + 7) * PX * 4 / 1e6, "U-Net output + cloth in, 3 images / masks out"),
    "l2norm_corr": (F * (2 * 192 * 512 * 4 + planes(192 * 192)) / 1e6, "two feature maps in + correlation planes out"),
}
FLOP = {"conv_igemm_kernel": 16.003 * F, "sagan_attention": F * 2 * (192 * 192 + 2 * 48 * 48 + 12 * 12) * 576 / 1e9,
        "l2norm_corr": F * 2 * 192 * 192 * 512 / 1e9}


def main(path, passes):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    h = rows[0]
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    agg = defaultdict(float)
    for r in rows[1:]:
        if r[mi] == "gpu__time_duration.sum":
            agg[r[ki]] += float(r[vi].replace(",", "")) / 1e3 / passes  # us per step
    cls = defaultdict(float)
    for k, us in agg.items():
        name = next((f for f in list(ALG) + list(FLOP) if f in k), None)
        cls[name or "other (weight packing in the first pass, torch glue)"] += us
    tot = sum(cls.values())
    print("| kernel class | us / step (ncu, cold) | share | algorithmic work / step | achieved | of measured peak |")
    print("|---|---:|---:|---|---:|---:|")
    for name, us in sorted(cls.items(), key=lambda kv: -kv[1]):
        work, ach, frac = "", "", ""
        if name in FLOP:
            gf = FLOP[name]
            mult = " issued (x3 MMAs in fp16x3)" if name == "conv_igemm_kernel" else " fp32"
            work, ach = f"{gf:.0f} GFLOP{mult}", f"{gf / us * 1e3:.0f} TFLOP/s"
            frac = f"{gf / us * 1e3 / TF:.2f} of tensor peak ({3 * gf / us * 1e3 / TF:.2f} as issued MMAs)" if name == "conv_igemm_kernel" else ""
        if name in ALG and name not in ("l2norm_corr",):
            mb, what = ALG[name]
            work, ach, frac = f"{mb:.0f} MB: {what}", f"{mb / us * 1e3 / 1e3:.2f} TB/s", f"{mb / us * 1e3 / HBM:.2f} of HBM"
        print(f"| `{name}` | {us:.0f} | {us / tot * 100:.1f}% | {work} | {ach} | {frac} |")
    print(f"| **total** | {tot:.0f} | 100% | | | |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
