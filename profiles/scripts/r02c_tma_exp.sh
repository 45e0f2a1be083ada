# experiment: cell-block fetch through per-thread bulk copies (QRMB_TMA=1) — parity first (short timeouts: a lost completion would spin), then A/B
mkdir -p gpurun_out
RLRM_LIB_PATH=$PWD/build/variants/librlrm_tma.so timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "float64_block or generic_kernel_equals" 2>&1 | tail -3
WL="cfg3_f64 ow_exp6_qrm cfg4_qrm" REP=1 TAG=r02_t39 timeout 400 bash profiles/scripts/ab_variants.sh
