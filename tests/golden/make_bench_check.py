"""Derives tests/golden/bench_check.json: what bench.py's parity probe must produce, computed with the CPU oracle.

bench.parity_probe runs one extra batch of the benchmark's own inputs through the timed VppRsgmPipeline with a fixed call
number and hashes the projected pair + the disparity bits of frames 0 and 63.  Here the same two frames go through the oracle:
the device generator's pattern stream restated on the host (vpp_core_opt.device_pattern), orc.vpp in numba arithmetic
(vpp_standalone.py:243-369), orc.compute_rsgm (models/rsgm/rsgm.py:250-294) with the recorded Intel RCPSS table that the
library uses by default.  No GPU involved.

    python tests/golden/make_bench_check.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import bench
    from oracle import oracle as orc
    from vppstereo_b200 import vpp_core_opt
    orc.build()
    lut = np.load(os.path.join(ROOT, "tests", "golden", "rcp_lut.npz"))["lut"]
    B, seed0 = 64, 1234                                              # bench batch; VppRsgmPipeline's default seed
    left, right, hints = bench.bench_inputs(0, B)
    rng_seed = (seed0 * 0x9E3779B97F4A7C15 + bench.PROBE_STEP) & (2**64 - 1)      # VppRsgmPipeline.pattern_seed
    digests = []
    for f in bench.PROBE_FRAMES:
        pattern = vpp_core_opt.device_pattern(rng_seed, f, orc.stream_length(hints[f], 3, 3, False))
        lw, rw = orc.vpp(left[f], right[f], hints[f], wsize=3, blending=0.4, stream=pattern, mode=1)
        disp = orc.compute_rsgm(left[f], lw, rw, dmax=bench.D, rcp_lut_override=lut)
        digests.append(bench.probe_digest(lw, rw, disp))
        print(f"frame {f}: mean disparity {disp.mean():.4f}  {digests[-1]}")
    out = {"what": "sha256(projected left | projected right | float32 disparity bits) of frames 0 and 63 of bench.py's parity probe",
           "made_by": "tests/golden/make_bench_check.py (CPU oracle, recorded Intel RCPSS table)",
           "batch": B, "pipeline_seed": seed0, "step": bench.PROBE_STEP, "frames": list(bench.PROBE_FRAMES), "sha256": digests}
    with open(bench.BENCH_CHECK, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", bench.BENCH_CHECK)


if __name__ == "__main__":
    main()
