"""Host builds of the device headers -- TEST-ONLY helpers (see tests/hostsim/*.cpp)."""
import ctypes
import os
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
BUILD = HERE / "_build"


def build(name: str) -> ctypes.CDLL:
    BUILD.mkdir(exist_ok=True)
    src = HERE / f"{name}.cpp"
    so = BUILD / f"lib{name}.so"
    deps = [src] + list((HERE.parent.parent / "pyjpegdecoder_b200" / "csrc").glob("*.cuh")) \
        + list((HERE.parent.parent / "include").glob("*.h"))
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        tmp = BUILD / f"lib{name}.{os.getpid()}.tmp.so"     # pytest-xdist workers may build at the same time
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++17",
                        "-o", str(tmp), str(src)], check=True)
        os.replace(tmp, so)
    return ctypes.CDLL(str(so))
