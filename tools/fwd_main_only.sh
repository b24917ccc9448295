#!/bin/bash
# time k_pool_fwd alone (ncu serialises the two forward kernels); usage: fwd_main_only.sh "ENV=.. ENV=.."
env $1 VEON_FWD_PDL=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pool_fwd -c 6 --csv python tools/fwd_ceiling.py C2 2>/dev/null | grep -E "k_pool_fwd<" | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
