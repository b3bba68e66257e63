#!/bin/bash
# round-2 pass O: the standard pass (tests, smoke, both bench arms, ncu launch list + full capture) + sanitizer pass 4
set -u
bash tools/gpu_round.sh
bash tools/sanitize4.sh
