#!/bin/bash
# SASS evidence for the hot kernels (run where cuobjdump is installed; no GPU needed):
#   tools/sass_excerpts.sh > profiles/r2_sass_excerpts.txt
lib=dolfinx_eqlb_b200/csrc/libeqlb_b200.so
echo "# cuobjdump -sass $lib (sm_100a), instruction counts per kernel; source hash $(python -c 'import bench; print(bench.kernel_source_hash())')"
cuobjdump -sass $lib | awk '
/Function :/ { name=$3; next }
name != "" {
  if ($0 ~ /REDG?\.E\.ADD\.F64/) red[name]++
  if ($0 ~ /MUFU\.RCP64H/) rcp[name]++
  if ($0 ~ /[^A-Z]STL/) stl[name]++
  if ($0 ~ /[^A-Z]LDL/) ldl[name]++
  if ($0 ~ /DFMA/) dfma[name]++
  if ($0 ~ /DMUL/) dmul[name]++
  if ($0 ~ /DADD/) dadd[name]++
  if ($0 ~ /SHFL/) shfl[name]++
  if ($0 ~ /LDS/) lds[name]++
  if ($0 ~ /LDG/) ldg[name]++
  if ($0 ~ /ACQBULK|UTMALDG|UBLKCP|LDGSTS/) asyncc[name]++
  if ($0 ~ /^ +\/\*[0-9a-f]+\*\/ /) tot[name]++
}
END {
  for (n in tot) printf "total=%-6d DFMA=%-5d DMUL=%-4d DADD=%-4d SHFL=%-4d LDS=%-4d LDG=%-3d REDG.E.ADD.F64=%-3d MUFU.RCP64H=%-3d STL=%-3d LDL=%-3d LDGSTS=%-3d %s\n", tot[n], dfma[n], dmul[n], dadd[n], shfl[n], lds[n], ldg[n], red[n], rcp[n], stl[n], ldl[n], asyncc[n], n
}' | c++filt | sed 's/(anonymous namespace):://; s/(PatchView.*//; s/(int, int const.*//; s/(eqlb.*//' | grep "patch_k2w_kernel\|patch_kw_kernel\|patch_k1w_kernel\|greedy_colour\|bc_poly\|halo_push" | sort -k13
