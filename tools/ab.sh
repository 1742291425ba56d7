#!/bin/bash
# A/B the kernel variants present in csrc/ with the short bench
cd "$(dirname "$0")/.."
for so in informative_path_planning_b200/csrc/libipp_b200*.so; do
  IPP_B200_LIB=$PWD/$so python bench.py --steps 3 --warmup 3 --no-extras | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$so'.split('/')[-1], 'q/s %.0f ms %.1f TF %.2f frac %.3f'%(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac']))"
done
