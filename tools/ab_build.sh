#!/bin/bash
# Build a tuning variant of the library from an alternative prefill_tcgen05.cu:  tools/ab_build.sh <name> <file.cu>
set -e
cd "$(dirname "$0")/.."
cp physics_llm_inference_b200/csrc/prefill_tcgen05.cu /tmp/_ab_saved.cu
cp "$2" physics_llm_inference_b200/csrc/prefill_tcgen05.cu
python -m physics_llm_inference_b200.build --variant="$1" --force | tail -1
cp /tmp/_ab_saved.cu physics_llm_inference_b200/csrc/prefill_tcgen05.cu
touch physics_llm_inference_b200/csrc/prefill_tcgen05.cu
