#!/bin/bash
O=gpurun_out; mkdir -p $O
GDN_SYNC_DEBUG=1 GDN_GRAPH=0 timeout 600 python tools/profile_step.py 1 > $O/debug_step.log 2>&1
tail -5 $O/debug_step.log
GDN_AUTOTUNE=0 GDN_SYNC_DEBUG=1 GDN_GRAPH=0 timeout 600 python tools/profile_step.py 1 > $O/debug_step_noat.log 2>&1
tail -5 $O/debug_step_noat.log
