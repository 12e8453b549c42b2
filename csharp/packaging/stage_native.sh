#!/bin/bash
# Builds libdracob200.so (nvcc, sm_100a) and stages it where Draco.B200.targets packs it from:
#   csharp/runtimes/linux-x64/native/libdracob200.so
set -e
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../.." && pwd)"
python "$root/draco_sharp_b200/build.py"
mkdir -p "$here/../runtimes/linux-x64/native"
cp "$root/draco_sharp_b200/libdracob200.so" "$here/../runtimes/linux-x64/native/"
echo "staged $(ls -la "$here/../runtimes/linux-x64/native/libdracob200.so")"
