#!/bin/bash
mkdir -p gpurun_out
MSIG_LIB=$PWD/multi-domain-style-injected-gan_b200/libmsig_prof.so timeout 300 python profiles/probe/ring_profile.py 2>&1 | tee gpurun_out/ring_profile.txt
