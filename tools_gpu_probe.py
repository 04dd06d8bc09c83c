"""Run groups of GPU tests in separate processes (a trapped kernel poisons its CUDA context) and log the tails.
Usage on the GPU box:  python tools_gpu_probe.py [group ...]   -> gpurun_out/probe.log"""
import os
import subprocess
import sys
import time

GROUPS = {
    "elem": ["tests/test_gpu_ops.py", "-k", "groupnorm or layernorm or time_embedding or cfg_ddim"],
    "gemm": ["tests/test_gpu_ops.py", "-k", "gemm_plain"],
    "gemm_epi": ["tests/test_gpu_ops.py", "-k", "gemm_bias or gemm_geglu"],
    "gemm_fused": ["tests/test_gpu_ops.py", "-k", "layernorm_fold or fused"],
    "conv": ["tests/test_gpu_ops.py", "-k", "conv3x3"],
    "attn64": ["tests/test_gpu_ops.py", "-k", "attention and (64 or 32)"],
    "attn": ["tests/test_gpu_ops.py", "-k", "attention and not (64 or 32)"],
    "golden": ["tests/test_golden.py"],
    "driver": ["tests/test_gpu_driver.py"],
    "edges": ["tests/test_gpu_edges.py"],
    "unet_tiny": ["tests/test_gpu_unet.py", "-k", "not sd15 and not sd21"],
    "unet_sd15": ["tests/test_gpu_unet.py", "-k", "sd15"],
    "unet_sd21": ["tests/test_gpu_unet.py", "-k", "sd21"],
    "clip": ["tests/test_gpu_clip.py"],
    "preprocess": ["tests/test_preprocess.py"],
    "vae_tiny": ["tests/test_gpu_vae.py", "-k", "not sd_vae"],
    "vae_sd": ["tests/test_gpu_vae.py", "-k", "sd_vae"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(GROUPS)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/probe.log", "a") as log:
        for n in names:
            t0 = time.time()
            cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-s", "--tb=line", "-p", "no:cacheprovider", *GROUPS[n]]
            try:
                r = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
                out, rc = r.stdout + r.stderr, r.returncode
            except subprocess.TimeoutExpired as e:
                out, rc = (e.stdout or b"").decode(errors="replace") + "\nTIMEOUT", -9
            tail = "\n".join(out.splitlines()[-60:])
            msg = f"===== {n}: rc={rc} {time.time() - t0:.1f}s =====\n{tail}\n"
            log.write(msg)
            log.flush()
            print(msg, flush=True)
