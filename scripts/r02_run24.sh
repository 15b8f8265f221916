#!/bin/bash
# end of round 2 (after the load-chain work): driver-style check, then the profile set of the shipped kernels
out=gpurun_out
bash scripts/r02_final_check.sh
bash scripts/profile_round.sh r02_final5 > $out/r02_final5_profile.log 2>&1; echo "== profile rc $?"; tail -2 $out/r02_final5_profile.log
