#!/bin/bash
# round-2 pass E: the standard pass (tests, smoke, both bench arms, ncu launch list + full capture) + sanitizer
set -u
bash tools/gpu_round.sh
SKIP=1 bash tools/sanitize3.sh
