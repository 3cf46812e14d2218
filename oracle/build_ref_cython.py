"""TEST INFRASTRUCTURE ONLY.  Builds the reference's own Cython host code (lib/nms/cpu_nms.pyx: cpu_nms, cpu_soft_nms;
lib/bbox/bbox.pyx: bbox_overlaps_cython, ignore_overlaps_cython) into oracle/_ref/ from the sources where they lie
under /root/reference.  The sources target Python 2 / numpy < 1.20; the tokens that no longer exist are rewritten ON THE
FLY into oracle/_ref/cy/ (git-ignored, nothing of the reference is committed):

    cpu_nms.pyx : `np.float thresh` -> `double thresh`   (np.float was the builtin float: a double)
                  `np.int_t`        -> `np.int32_t`      (the arrays are created with .astype('i') / dtype int)
                  `dtype=np.int`    -> `dtype=np.int32`
    bbox.pyx    : `DTYPE = np.float` -> `DTYPE = float`

Everything else, including the arithmetic, is the reference's.  Usage: python oracle/build_ref_cython.py [/root/reference]"""
import os
import re
import sys


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "_ref")
    cy = os.path.join(out, "cy")
    src_nms = os.path.join(ref, "lib", "nms", "cpu_nms.pyx")
    src_bbox = os.path.join(ref, "lib", "bbox", "bbox.pyx")
    if not (os.path.exists(src_nms) and os.path.exists(src_bbox)):
        print("reference tree not present: keeping prebuilt oracle/_ref")
        return 0
    os.makedirs(cy, exist_ok=True)
    s = open(src_nms).read()
    s = s.replace("np.float thresh", "double thresh").replace("np.int_t", "np.int32_t")
    s = re.sub(r"dtype=np\.int\b", "dtype=np.int32", s)
    open(os.path.join(cy, "ref_cpu_nms.pyx"), "w").write(s)
    s = open(src_bbox).read().replace("DTYPE = np.float\n", "DTYPE = float\n")
    open(os.path.join(cy, "ref_bbox.pyx"), "w").write(s)
    import numpy as np
    from Cython.Build import cythonize
    from setuptools import Extension, setup
    exts = [Extension("ref_cpu_nms", [os.path.join(cy, "ref_cpu_nms.pyx")], include_dirs=[np.get_include()],
                      extra_compile_args=["-O2", "-w"], define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")]),
            Extension("ref_bbox", [os.path.join(cy, "ref_bbox.pyx")], include_dirs=[np.get_include()],
                      extra_compile_args=["-O2", "-w"], define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")])]
    os.environ.setdefault("CC", "/usr/bin/gcc")
    setup(name="ref_cython", ext_modules=cythonize(exts, language_level=2, quiet=True),
          script_args=["-q", "build_ext", "--build-lib", out, "--build-temp", os.path.join(cy, "build")])
    return 0


if __name__ == "__main__":
    sys.exit(main())
