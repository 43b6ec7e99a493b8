#!/bin/bash
# ncu --set full of one k_bwd_mlp launch of the zero-pad training iteration (after the plain run has exited 0)
ZP_ONLY=1 python scripts/time_zeropad_train.py > gpurun_out/zp_plain.log 2>&1 && \
ZP_ONLY=1 ncu --set full --clock-control none --import-source on -k regex:k_bwd_mlp --launch-skip 120 -c 1 -f -o gpurun_out/r02_k_bwd_mlp python scripts/time_zeropad_train.py > gpurun_out/zp_ncu.log 2>&1
tail -2 gpurun_out/zp_ncu.log; ls -la gpurun_out/r02_k_bwd_mlp.ncu-rep
