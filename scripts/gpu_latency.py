"""Latency of a warp-wide load of 32 distinct random sectors (a dependent chain per lane) against the number of resident
warps per SM, over buffers of the size of K1's index (L2 resident) -- kbo_measure_random_sector_rate(dependent = warps)."""
import json
from kbo_b200 import api, build
build.build_library()
api.load_library()
out = {}
for mb in (46, 64, 80):
    for warps in (4, 8, 16, 24, 32, 40, 48, 64):
        rate = api.measure_random_sector_rate(mb << 20, warps, 0)
        lanes = 148 * warps * 32
        ns = lanes / rate * 1e9
        out["%dMB_%dwarps" % (mb, warps)] = {"sectors_per_s": rate, "ns_per_load": ns, "cycles_per_load_at_1965MHz": ns * 1.965}
        print(mb, "MB", warps, "warps/SM: %.1f G sectors/s, %.0f ns = %.0f cycles per dependent load" % (rate / 1e9, ns, ns * 1.965), flush=True)
json.dump(out, open("gpurun_out/r2_latency_vs_occupancy.json", "w"), indent=1)
