"""Build recipe for oracle/_ref: the reference's OWN CUDA ops, compiled where they lie.

TEST INFRASTRUCTURE ONLY.  Compiles /root/reference/op/{upfirdn2d,fused_bias_act}{.cpp,_kernel.cu}
(reference `op/upfirdn2d.py:9-16`, `op/fused_act.py:10-17` JIT-build the same four files) for sm_100a into
`oracle/_ref/` as two torch extensions (`upfirdn2d_ref.so`, `fused_ref.so`).  No source is copied into the repo:
the compiler reads the files from /root/reference.  `oracle/_ref/` is git-ignored but travels to the GPU box with
gpurun, where `-m gpu` tests use it as the bit-level comparator for our upfirdn2d / fused_bias_act kernels.

/root/reference does not exist on the GPU box: there this script is a no-op and `load_ref()` just imports the
prebuilt .so files.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = "/root/reference/op"


def build(verbose=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    done = True
    for name, srcs in (
        ("upfirdn2d_ref", ["upfirdn2d.cpp", "upfirdn2d_kernel.cu"]),
        ("fused_ref", ["fused_bias_act.cpp", "fused_bias_act_kernel.cu"]),
    ):
        if os.path.exists(os.path.join(OUT, name + ".so")):
            continue
        try:
            load(
                name,
                sources=[os.path.join(REF, s) for s in srcs],
                build_directory=OUT,
                extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"],
                verbose=verbose,
                is_python_module=True,
            )
        except Exception as e:  # the checker is optional; never break build() over it
            print(f"[oracle/_ref] build of {name} failed: {e}", file=sys.stderr)
            done = False
    return done


def load_ref(name):
    """Import a prebuilt reference extension ("upfirdn2d_ref" / "fused_ref") or return None."""
    path = os.path.join(OUT, name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (libtorch symbols must be loaded first)

    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print("built" if build(verbose=True) else "skipped/failed")
