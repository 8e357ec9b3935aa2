#!/bin/bash
# developer tool: time the config-2 sweep with alternative builds of the library (variants_tmp/*.so)
cp mcpower_b200/libmcpower_b200.so /tmp/lib_orig.so
for f in /tmp/lib_orig.so variants_tmp/*.so; do
  cp $f mcpower_b200/libmcpower_b200.so
  echo -n "$f: "; python tools/run_c2_sweep.py
done
cp /tmp/lib_orig.so mcpower_b200/libmcpower_b200.so
