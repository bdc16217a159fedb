import os, sys, subprocess, json, ctypes as C
sys.path.insert(0, '/root/repo')
if len(sys.argv) > 1 and sys.argv[1] == 'probe':
    from commander_b200 import sharp
    L = sharp.lib()
    L.cmdr_sht_measure_dmma_tflops.argtypes = [C.c_int, C.c_int, C.c_int]; L.cmdr_sht_measure_dmma_tflops.restype = C.c_double
    print("DFMA only TF:", sharp.measure_fp64_tflops(4096, 5))
    print("DMMA only TF:", L.cmdr_sht_measure_dmma_tflops(2048, 5, 0))
    print("DMMA+DFMA TF:", L.cmdr_sht_measure_dmma_tflops(2048, 5, 1))
    sys.exit(0)
for var, vals in (("CMDR_SHT_R_A2", (2, 3, 4)), ("CMDR_SHT_R_S2", (2, 3, 4)), ("CMDR_SHT_R_A0", (4, 6, 8)), ("CMDR_SHT_R_S0", (4, 6, 8))):
    for v in vals:
        env = dict(os.environ); env[var] = str(v)
        out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--e2e-steps", "1"],
                             env=env, capture_output=True, text=True, cwd='/root/repo').stdout.strip().splitlines()
        try:
            j = json.loads(out[-1]); k = j["roofline"]["kernels"]
            print(var, v, "pairs/s %.2f" % j["value"], {a: b for a, b in k.items() if a.endswith("_ms")}, flush=True)
        except Exception as e:
            print(var, v, "FAILED", out[-3:] if out else e, flush=True)
