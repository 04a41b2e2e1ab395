#!/usr/bin/env bash
TAG=${1:-x}
bash tools/gpu_ncu.sh $TAG enc.2.r3:fwd enc.0.skip:fwd enc.3.r3:wgrad
