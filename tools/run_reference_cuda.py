"""Run the reference's own CUDA trainer (oracle/_ref/train_unet[_log10], built by oracle/Makefile) for a bounded time on
synthetic data and print its log: python tools/run_reference_cuda.py [seconds] [ncu_csv skip count]
With an ncu_csv argument the trainer runs under `ncu --metrics gpu__time_duration.sum` for `count` launches after
`skip` (device time of every kernel of a few steps: is the step GPU-bound or host-bound?)."""
import os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import unet_oracle as O
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 40
exe = os.path.join(ROOT, "oracle", "_ref", "train_unet")
if os.path.exists(exe + "_log10"):
    exe += "_log10"
ncu_csv = os.path.abspath(sys.argv[2]) if len(sys.argv) > 2 else None
pre = []
if ncu_csv:
    pre = ["ncu", "--metrics", "gpu__time_duration.sum", "--clock-control", "none", "-s", sys.argv[3], "-c", sys.argv[4],
           "--csv", "--log-file", ncu_csv]
with tempfile.TemporaryDirectory() as d:
    cfg = O.UNetConfig()
    flat = O.flatten_params(cfg, O.init_params(cfg, seed=0)).numpy()
    O.write_model_bin(os.path.join(d, "unet_init.bin"), cfg, flat, B=32)
    O.write_data_bin(os.path.join(d, "data.bin"), np.random.default_rng(0).uniform(-1, 1, (256, 3, 64, 64)).astype(np.float32))
    os.makedirs(os.path.join(d, "data"), exist_ok=True)
    t0 = time.time()
    p = subprocess.Popen(pre + [exe, "--model_weights", "unet_init.bin", "--data_file", "data.bin", "--log_file", "log.txt"],
                         cwd=d, stdout=open(os.path.join(d, "out.txt"), "w"), stderr=subprocess.STDOUT, env=dict(os.environ, CUDA_MODULE_LOADING="EAGER"))
    while time.time() - t0 < secs and p.poll() is None:
        time.sleep(1)
    print("returncode", p.poll(), "after", time.time() - t0)
    if pre and p.poll() is None:
        # under ncu: end the TRAINER (ncu's child, by pid), so that ncu exits by itself and writes its log file
        kids = subprocess.run(["ps", "-o", "pid=", "--ppid", str(p.pid)], stdout=subprocess.PIPE, text=True).stdout.split()
        for k in kids:
            os.kill(int(k), 15)
        try:
            p.wait(timeout=60)
        except subprocess.TimeoutExpired:
            pass
    if p.poll() is None:
        p.kill()
    p.wait()
    for f in ("out.txt", "log.txt"):
        fp = os.path.join(d, f)
        print("----", f)
        if os.path.exists(fp):
            print(open(fp).read()[-2500:])
