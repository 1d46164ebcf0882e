#!/bin/bash
# tools/devignore.sh on|off — keep the 270 MB of shipped weights out of DEV gpurun pushes (5 GPU-minutes each). MUST be `off` at round end:
# the driver's GPU tests need oracle/_ref/models/.
if [ "$1" = "on" ]; then
  grep -q "DEV ONLY" .gpurunignore || printf '# DEV ONLY (remove before round end): shipped weights cost 5 GPU-minutes of push per call\noracle/_ref/models\noracle/_ref/models/*\n' >> .gpurunignore
else
  sed -i '/DEV ONLY/,+2d' .gpurunignore
fi
cat .gpurunignore
