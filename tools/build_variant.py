"""Developer tool: build libmilb200 with constants of csrc/tc_gemm.cu overridden, into gpurun_out-free scratch
(llm-guided-multimodal-mil_b200/_variants/<name>.so, git-ignored), for A/B runs with MILB200_LIB=<path>.
Usage: python tools/build_variant.py name TNG_ASTAGES=6 TNG_BSTAGES=5 ..."""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "llm-guided-multimodal-mil_b200")
sys.path.insert(0, PKG)
import build as B  # noqa: E402


def main():
    name, sets = sys.argv[1], dict(a.split("=") for a in sys.argv[2:])
    out_dir = os.path.join(PKG, "_variants")
    os.makedirs(out_dir, exist_ok=True)
    src = open(os.path.join(PKG, "csrc", "tc_gemm.cu")).read()
    for k, v in sets.items():
        src, n = re.subn(r"(constexpr int %s = )[^;]+;" % re.escape(k), r"\g<1>%s;" % v, src, count=1)
        assert n == 1, k
    tmp = os.path.join(PKG, "csrc", "_variant_%s_tc_gemm.cu" % name)
    open(tmp, "w").write(src)
    try:
        obj = os.path.join(out_dir, name + "_tc_gemm.o")
        subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-c", tmp, "-o", obj], check=True)
    finally:
        os.remove(tmp)
    objs = [os.path.join(B.OBJ, s[:-3] + ".o") for s in B.sources() if s != "tc_gemm.cu" and not s.startswith("_variant")] + [obj]
    lib = os.path.join(out_dir, name + ".so")
    subprocess.run([B._nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", lib] + objs + ["-lcudart"], check=True)
    os.remove(obj)
    print(lib)


if __name__ == "__main__":
    main()
