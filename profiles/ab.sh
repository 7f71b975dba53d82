#!/bin/bash
# A/B timing of alternative builds of libgd_b200.so on the SAME box (box-to-box variation is 2-4 %):
#   profiles/ab.sh "<conv_sweep args>" libA.so libB.so ...   (libraries under guided_diffusion_clip_b200/build/ab/)
args="$1"; shift
for rep in 1 2; do
  for lib in "$@"; do
    echo "== $lib (pass $rep)"
    eval "GD_B200_LIB=$PWD/guided_diffusion_clip_b200/build/ab/$lib python profiles/conv_sweep.py $args"
  done
done
