"""Run the reference's own CUDA trainer (oracle/_ref/train_unet, built by oracle/Makefile) for a bounded time on
synthetic data and print its log: python tools/run_reference_cuda.py [seconds]"""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unet_oracle as O
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 40
exe = os.path.join(ROOT, "oracle", "_ref", "train_unet")
with tempfile.TemporaryDirectory() as d:
    cfg = O.UNetConfig()
    flat = O.flatten_params(cfg, O.init_params(cfg, seed=0)).numpy()
    O.write_model_bin(os.path.join(d, "unet_init.bin"), cfg, flat, B=32)
    O.write_data_bin(os.path.join(d, "data.bin"), np.random.default_rng(0).uniform(-1, 1, (256, 3, 64, 64)).astype(np.float32))
    os.makedirs(os.path.join(d, "data"), exist_ok=True)
    t0 = time.time()
    p = subprocess.Popen([exe, "--model_weights", "unet_init.bin", "--data_file", "data.bin", "--log_file", "log.txt"],
                         cwd=d, stdout=open(os.path.join(d, "out.txt"), "w"), stderr=subprocess.STDOUT, env=dict(os.environ, CUDA_MODULE_LOADING="EAGER"))
    while time.time() - t0 < secs and p.poll() is None:
        time.sleep(1)
    print("returncode", p.poll(), "after", time.time() - t0)
    p.kill(); p.wait()
    for f in ("out.txt", "log.txt"):
        fp = os.path.join(d, f)
        print("----", f)
        if os.path.exists(fp):
            print(open(fp).read()[-2500:])
