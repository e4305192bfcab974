"""Builds the pybind11 host mirror (VegasAfterglowC_b200) against libvag_b200.so."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)


def main():
    import numpy
    import pybind11

    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(PKG, "VegasAfterglowC_b200" + ext)
    src = os.path.join(HERE, "vag_pybind.cpp")
    deps = [src, os.path.join(PKG, "..", "include", "vag.h"), os.path.join(PKG, "libvag_b200.so")]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", src, "-o", out,
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(),
           "-L" + PKG, "-lvag_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)


if __name__ == "__main__":
    main()
