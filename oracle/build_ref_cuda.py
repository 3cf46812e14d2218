"""ORACLE recipe -- test infrastructure only.

Compiles the REFERENCE's own CUDA kernels for this path into oracle/_ref/libref_cuda{,_fma}.so so that the GPU tests can
run them on the B200 next to the product kernels and the C oracle (they pin the rules that the reference implements on
the GPU only):
  SNIPER-mxnet/src/operator/multi_proposal_target.cu       :75-114  GenerateAnchors (host)
                                                           :116-331 NonMaximumSuppression, getProps (kernels)
                                                           :435-578 GT append + IoU / label / target assignment (host)
  SNIPER-mxnet/src/operator/contrib/deformable_psroi_pooling.cu :48-161, 202-330   bilinear_interp, fwd, bwd kernels
  SNIPER-mxnet/src/operator/contrib/psroi_pooling.cu            :50-118, 145-212   fwd, bwd kernels
  SNIPER-mxnet/src/operator/contrib/nn/deformable_im2col.cuh    :77-262, 316-364, 418-473   im2col / col2im / coord
Nothing is copied into the repository: the line ranges are cut out of the reference tree ON THE FLY into the git-ignored
oracle/_ref/cu/*.inc (each cut is checked against the text it must start and end with), and oracle/ref_cuda_harness.cu
(ours: a prelude for the four mshadow/nnvm names the cuts use, and extern "C" launchers over raw device pointers)
#includes them.  Two builds: `libref_cuda_fma.so` with nvcc's defaults (what the reference's own Makefile produces: FMA
contraction on) and `libref_cuda.so` with -fmad=false (the C abstract machine the oracle restates); the tests compare
bit-exactly against the latter and report the deltas of the former.
Usage: python oracle/build_ref_cuda.py [/root/reference]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
CU = os.path.join(OUT, "cu")

# (file under SNIPER-mxnet/src/operator, first line, last line, must-start-with, must-end-with, output name)
CUTS = [
    ("multi_proposal_target.cu", 75, 114, "inline void _MakeAnchor", "}", "mpt_anchors.inc"),
    ("multi_proposal_target.cu", 116, 331, "// greedily keep the max detections", "}", "mpt_kernels.inc"),
    ("multi_proposal_target.cu", 435, 578, "std::vector <int> numgts_per_image(num_images);", "}", "mpt_host_assign.inc"),
    ("contrib/deformable_psroi_pooling.cu", 48, 161, "template <typename DType>", "}", "dpsroi_fwd.inc"),
    ("contrib/deformable_psroi_pooling.cu", 202, 330, "template <typename DType>", "}", "dpsroi_bwd.inc"),
    ("contrib/psroi_pooling.cu", 50, 118, "template <typename DType>", "}", "psroi_fwd.inc"),
    ("contrib/psroi_pooling.cu", 145, 212, "template <typename DType>", "}", "psroi_bwd.inc"),
    ("contrib/nn/deformable_im2col.cuh", 77, 262, "template <typename DType>", "}", "dim2col_fwd.inc"),
    ("contrib/nn/deformable_im2col.cuh", 316, 364, "template <typename DType>", "}", "dim2col_col2im.inc"),
    ("contrib/nn/deformable_im2col.cuh", 418, 473, "template <typename DType>", "}", "dim2col_coord.inc"),
]


def cut(ref):
    os.makedirs(CU, exist_ok=True)
    base = os.path.join(ref, "SNIPER-mxnet", "src", "operator")
    for rel, a, b, start, end, name in CUTS:
        lines = open(os.path.join(base, rel)).read().split("\n")
        seg = lines[a - 1:b]
        if not seg[0].strip().startswith(start) or seg[-1].strip() != end:
            raise RuntimeError("%s:%d-%d is not the expected text (%r ... %r)" % (rel, a, b, seg[0], seg[-1]))
        with open(os.path.join(CU, name), "w") as f:
            f.write("// cut at build time from %s:%d-%d (reference tree; not part of this repository)\n" % (rel, a, b))
            f.write("\n".join(seg) + "\n")


def build(ref="/root/reference"):
    if not os.path.isdir(os.path.join(ref, "SNIPER-mxnet")):
        print("reference tree not present: keeping prebuilt oracle/_ref/libref_cuda*.so")
        return False
    cut(ref)
    src = os.path.join(HERE, "ref_cuda_harness.cu")
    common = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++14", "-shared", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fopenmp", "-w", "-I", CU, src]
    subprocess.check_call(common + ["-fmad=false", "-o", os.path.join(OUT, "libref_cuda.so")])
    subprocess.check_call(common + ["-o", os.path.join(OUT, "libref_cuda_fma.so")])
    return True


if __name__ == "__main__":
    print(build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
