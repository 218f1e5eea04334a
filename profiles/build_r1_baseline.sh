#!/bin/bash
# Round-1 library as an A/B baseline: the csrc tree of the round-1 commit (302c369), built with the same flags into
# marl-mass_b200/_build/variants/lib_r1.so, plus a stub for the entry point added in round 2 so that the round-2 Python
# layer can load it (MM_LIB_PATH).  Run here (nvcc cross-compiles); the .so travels to the GPU box with gpurun.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" archive 302c369 marl-mass_b200/csrc include | tar -x -C "$TMP"
# round 2 prepended `int32_t struct_size` to mm_config (ABI guard): give the round-1 header the same layout, and stubs
# for the entry points round 2 added, so that the round-2 binding drives the round-1 kernels unchanged
sed -i 's|    int32_t shield;          /\* safety_guarantee|    int32_t struct_size;\n    int32_t shield;          /* safety_guarantee|' "$TMP/include/marl_mass_b200.h"
grep -q "int32_t struct_size" "$TMP/include/marl_mass_b200.h"
cat > "$TMP/stub.cu" <<'EOS'
extern "C" int mm_step_build(const void *) { return 0; }
extern "C" int mm_abi_version(void) { return 3; }
EOS
mkdir -p "$ROOT/marl-mass_b200/_build/variants"
cd "$TMP/marl-mass_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -shared \
  -I "$TMP/include" -o "$ROOT/marl-mass_b200/_build/variants/lib_r1.so" \
  merge_step.cu merge_step_occ4.cu actor_sample.cu supervisor.cu capi.cu "$TMP/stub.cu"
rm -rf "$TMP"
echo built lib_r1.so
