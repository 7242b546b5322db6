"""CPU: static invariants of the CUDA sources that no GPU test can catch reliably.

Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start before
its predecessor in the stream has finished, so it MUST execute griddepcontrol.wait (pdl_wait) before it touches global
memory.  A kernel that forgets it reads stale data only when the timing is unlucky -- so the rule is enforced on the
source: every kernel handed to mtl_launch_pdl (or launched with the attribute through cudaLaunchKernelEx in gemm_tc.cu)
calls pdl_wait()."""
import glob
import os
import re

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "meta-transfer-learning_b200", "csrc")


def _sources():
    return {os.path.basename(f): open(f).read() for f in glob.glob(os.path.join(CSRC, "*.cu"))}


def _kernel_body(src, name):
    m = re.search(r"(__global__|__device__ __forceinline__ void)[^;{]*?\b" + re.escape(name) + r"\s*\(", src)
    if not m:
        return None
    rest = src[m.end():]
    nxt = rest.find("__global__")
    body = rest if nxt < 0 else rest[:nxt]
    # a kernel that is only a named wrapper around a shared device body (gemm_tc_kernel / conv_gemm_tc_kernel):
    # the rule applies to the body it calls
    call = re.search(r"\b([A-Za-z0-9_]+_body)\s*<", body[:600])
    if call and "pdl_wait()" not in body[:600]:
        return _kernel_body(src, call.group(1))
    return body


def test_every_pdl_launched_kernel_waits_before_touching_memory():
    src = _sources()
    names = set()
    for s in src.values():
        names.update(re.findall(r"mtl_launch_pdl\(\s*([A-Za-z0-9_]+)", s))
    names.discard("kern")                                   # template launchers of gemm_tc.cu: listed explicitly below
    names.update({"gemm_tc_kernel", "conv_gemm_tc_kernel", "conv3x3_kw_kernel", "conv3x3_wgrad_kw_kernel"})
    assert len(names) >= 10, names
    for n in sorted(names):
        bodies = [b for b in (_kernel_body(s, n) for s in src.values()) if b is not None]
        assert bodies, f"kernel {n} not found"
        assert all("pdl_wait()" in b for b in bodies), f"{n} is launched with the PDL attribute but never calls pdl_wait()"


def test_product_sources_never_reference_the_oracle():
    pkg = os.path.dirname(CSRC)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(root, f)
