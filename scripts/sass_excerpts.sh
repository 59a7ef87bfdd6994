#!/bin/bash
# Writes profiles/r02_sass_*.txt: the SASS of the kernels the bench launches (cuobjdump -sass), cut to the
# instruction text, plus a count of the mnemonics that prove the TMA / mbarrier / PDL paths are real.
B=bloomsearch_b200/_build
out=profiles
dump() { # object, function-name regex, output file
  fn=$(cuobjdump -sass "$1" | grep -oE "Function : [A-Za-z0-9_]+" | awk '{print $3}' | grep -E "$2" | head -1)
  { echo "# cuobjdump -sass -fun $fn $1  (sm_100a, nvcc 12.9, -O3 -lineinfo)"; cuobjdump -sass -fun "$fn" "$1" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/[[:space:]]+$//'; } > "$3"
  echo "$3: $(wc -l < "$3") instructions; UBLKCP $(grep -c UBLKCP "$3"), SYNCS $(grep -c 'SYNCS' "$3"), ATOMS $(grep -c ATOMS "$3"), RED $(grep -c 'RED\.' "$3"), IMAD.HI $(grep -c 'IMAD.HI' "$3"), VOTE $(grep -c VOTE "$3"), BAR.SYNC $(grep -c 'BAR.SYNC' "$3")"
}
dump $B/kernels_probe_tiles.cu.o 'probe_tiles_kernelILi3ELi1024ELb0' $out/r02_sass_probe_tiles_nt3_1024.txt
dump $B/kernels_probe.cu.o 'probe_staged2_kernelILi16ELi2ELi3ELi16ELi16ELb0' $out/r02_sass_probe_staged2_16_2_3_16_16.txt
dump $B/kernels_probe.cu.o 'probe_staged2_kernelILi16ELi2ELi3ELi16ELi4ELb0' $out/r02_sass_probe_staged2_16_2_3_16_4.txt
dump $B/kernels_build.cu.o 'bsg12build_kernel' $out/r02_sass_build_kernel.txt
dump $B/bsg_comm.cpp.o 'or_reduce_p2p_kernel' $out/r02_sass_or_reduce_p2p.txt
