#!/bin/bash
# host-fed pipeline: grid caps of the per-slice reach-box kernels (ARTP_PIPE_CAPS="g,f", units of half the SM count; 0 = none)
: > gpurun_out/pt.log
for pc in ${CAPS:-0,0 4,4 4,5 5,4 3,4 4,3 5,5 3,3 4,4}; do echo "== ARTP_PIPE_CAPS=$pc" >> gpurun_out/pt.log; ARTP_PIPE_CAPS=$pc python profiles/e2e_probe.py 2>&1 | grep slice >> gpurun_out/pt.log; done
