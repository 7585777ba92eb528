#!/bin/bash
cd /root/repo
for w in 8 16 5; do echo "== long stress v3 W=$w"; timeout -s KILL 400 python tools/v3_stress.py $w 80 0 3 8 1 5 2>&1 | tail -2; done
