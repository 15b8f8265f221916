#!/bin/bash
# final state of round 2: full GPU suite + smoke + both bench arms, then the profile set (scripts/profile_round.sh)
out=gpurun_out
bash scripts/r02_final_check.sh
bash scripts/profile_round.sh r02_final2 > $out/r02_final2_profile.log 2>&1; echo "== profile rc $?"; tail -3 $out/r02_final2_profile.log
