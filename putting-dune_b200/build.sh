#!/usr/bin/env bash
# Builds libpdune_b200.so (sm_100a) in-tree. Parity-critical translation units
# (float64 geometry -> rate -> clock) are compiled with -fmad=false so that
# every product and sum is rounded exactly like the reference's NumPy code.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
src="$here/csrc"
out="$here/lib"
obj="${PD_OBJ_DIR:-$here/build}"
libname="${PD_LIB_NAME:-libpdune_b200.so}"
mkdir -p "$out" "$obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
COMMON=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17
        -Xcompiler -fPIC -I"$here/../include" -I"$src" ${PD_NVCC_EXTRA:-})
exact=(pd_lattice pd_reset pd_step pd_step_fast pd_query)
fast=(pd_api pd_mlp)
[ -f "$src/pd_render.cu" ] && fast+=(pd_render)
[ -f "$src/pd_synth.cu" ] && fast+=(pd_synth)
[ -f "$src/pd_render_cluster.cu" ] && fast+=(pd_render_cluster)
[ -f "$src/pd_episode.cu" ] && exact+=(pd_episode)
[ -f "$src/pd_env.cu" ] && exact+=(pd_env)
[ -f "$src/pd_mask.cu" ] && exact+=(pd_mask)
[ -f "$src/pd_export.cu" ] && exact+=(pd_export)
pids=()
for f in "${exact[@]}"; do
  "$NVCC" "${COMMON[@]}" -fmad=false -c "$src/$f.cu" -o "$obj/$f.o" & pids+=($!)
done
for f in "${fast[@]}"; do
  "$NVCC" "${COMMON[@]}" -c "$src/$f.cu" -o "$obj/$f.o" & pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
objs=()
for f in "${exact[@]}" "${fast[@]}"; do objs+=("$obj/$f.o"); done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out/$libname" "${objs[@]}" -lcudart
echo "built $out/$libname"
