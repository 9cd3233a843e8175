#!/bin/sh
# Cuts the //@emul-begin ... //@emul-end regions out of csrc/dis.cu and builds the host emulation next to this script.
set -e
here="$(cd "$(dirname "$0")" && pwd)"
out="${1:-$here/_build}"
mkdir -p "$out"
awk '/\/\/@emul-begin/{on=1; next} /\/\/@emul-end/{on=0} on' "$here/../../comfyui-video-stabilizer_b200/csrc/dis.cu" > "$out/vr_regions.inc"
g++ -O1 -std=c++17 -pthread -ffp-contract=off -Wno-unknown-pragmas -I"$out" "$here/vr_emul.cpp" -o "$out/vr_emul"
