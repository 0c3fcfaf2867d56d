"""Build the UNMODIFIED reference application (src/main.cpp + server / sink / source files)
against the drop-in headers and libraries. Needs /root/reference; the
binary lands in build/refapp/ (git-ignored, travels to the GPU box with the snapshot)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FMTUNER_REFERENCE", "/root/reference")
PKG = os.path.join(ROOT, "fmtuner_sdr_b200")
OUT = os.path.join(ROOT, "build", "refapp")
SOURCES = ["main", "config", "xdr_server", "audio_output", "rtl_sdr_device", "rtl_tcp_client",
           "signal_level", "cpu_features"]


def build(syntax_only: bool = False) -> str | None:
    if not os.path.isdir(os.path.join(REF, "src")):
        return None
    os.makedirs(OUT, exist_ok=True)
    objs = []
    for f in SOURCES:
        cmd = ["g++", "-std=c++17", "-DFM_SDR_TUNER_VERSION=\"1.3.0-b200\"", "-I",
               os.path.join(PKG, "dropin"), "-I", os.path.join(REF, "include"),
               os.path.join(REF, "src", f + ".cpp")]
        if syntax_only:
            subprocess.run(cmd + ["-fsyntax-only"], check=True)
            continue
        obj = os.path.join(OUT, f + ".o")
        subprocess.run(cmd + ["-O2", "-c", "-o", obj], check=True)
        objs.append(obj)
    if syntax_only:
        return ""
    exe = os.path.join(OUT, "fm-sdr-tuner-b200")
    subprocess.run(["g++", "-o", exe] + objs + ["-L", PKG, "-lfmgpu_dropin", "-lfmgpu", "-lssl",
                                                 "-lcrypto", "-lpthread", f"-Wl,-rpath,{PKG}"], check=True)
    return exe


if __name__ == "__main__":
    print(build("--syntax-only" in sys.argv))
