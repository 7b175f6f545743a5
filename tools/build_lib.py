"""Build exemplar_vae_b200/csrc/libexvae_b200.so without importing the package (whose import
requires an up-to-date library):   python tools/build_lib.py [--force] [-v]"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_build_module():
    spec = importlib.util.spec_from_file_location("exvae_build", os.path.join(ROOT, "exemplar_vae_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(load_build_module().build(force="--force" in sys.argv, verbose="-v" in sys.argv))
