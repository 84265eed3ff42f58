#!/bin/bash
# Build library variants for same-box A/B runs (scripts/ab_run.py):  scripts/ab.sh name1="-DX=1 ..." name2="..."
# Each variant is the working tree compiled with the given defines -> ab/librt_<name>.so.  `prev` = last commit.
set -e
cd "$(dirname "$0")/.."
mkdir -p ab; rm -f ab/*.so
C=parallel_ray_tracer_b200/csrc
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  if [ "$name" = "prev" ]; then
    git stash -q; rm -rf $C/build; make -C $C -j8 > /dev/null 2>&1; cp parallel_ray_tracer_b200/librt_b200.so ab/librt_prev.so; git stash pop -q
  else
    rm -rf $C/build; make -C $C -j8 RT_DEFS="$defs" > /dev/null 2>&1; cp parallel_ray_tracer_b200/librt_b200.so ab/librt_$name.so
  fi
done
rm -rf $C/build; make -C $C -j8 > /dev/null 2>&1
ls ab/
