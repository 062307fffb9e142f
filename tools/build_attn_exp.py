"""Build tools/bin/libattn_exp.so: the attention kernel source of the product compiled with -DS2V_ATTN_EXPERIMENT, which adds
the measurement-only entry point s2v_attn_fwd_exp (every warp-numbering / K-V-multicast / polynomial-fraction combination of the
same kernel, plus the per-CTA cycle counters).  The product library libs2v_b200.so is built WITHOUT that macro and exports none
of it.  Run here (nvcc cross-compiles); the .so travels to the GPU box with the snapshot."""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("disentangled-subject-to-vid_b200._build")
OUT = os.path.join(ROOT, "tools", "bin", "libattn_exp.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    objs = []
    for src_dir, name, extra in ((b.CSRC, "attn_tcgen05", ["-DS2V_ATTN_EXPERIMENT"]), (b.CSRC, "host_util", []),
                                 (os.path.join(ROOT, "tools"), "clock_probe", [])):
        obj = os.path.join(ROOT, "tools", "bin", name + "_exp.o")
        subprocess.run([b._nvcc(), *b.NVCC_FLAGS, *extra, "-c", os.path.join(src_dir, name + ".cu"), "-o", obj], check=True)
        objs.append(obj)
    subprocess.run([b._nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", OUT, *objs, "-lcudart"], check=True)
    return OUT


if __name__ == "__main__":
    print(build())
