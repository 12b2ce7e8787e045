#!/bin/bash
# round 2 session 2, call 26: force ms per launch / list ms per rebuild over the course of a run, per 20-step chunk: single domain vs two
# bricks (one GPU), bricks unsorted and sorted
cd "$(dirname "$0")/.."
python profiles/dd_case.py --nx 128 --bricks 2,1,1 --chunks 8 --opt sort_atoms=0 2>&1 | tail -2
python profiles/dd_case.py --nx 128 --bricks 2,1,1 --chunks 8 2>&1 | tail -1
