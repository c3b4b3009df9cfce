#!/bin/bash
# usage: tools/kb_build.sh <name> [extra nvcc -D flags...]   -> build/kb/<name>
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I zk-research-implementations_b200/csrc -I include "$@" tools/kbench.cu -o build/kb/$name
