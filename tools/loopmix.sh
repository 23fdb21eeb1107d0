#!/bin/bash
# Loop opcode mix (tools/sass_loop.py) of the production single-pass kernels in an object directory.
#   tools/loopmix.sh [objdir]     default: ckfft_b200/build
D=${1:-ckfft_b200/build}
show() {  # object, mangled-name regex, label
  f=$(cuobjdump -sass $D/$1 | grep "Function :" | sed 's/.*Function : //' | grep -E "$2" | head -1)
  [ -z "$f" ] && { echo "$3: not found"; return; }
  echo "== $3"; cuobjdump -sass -fun "$f" $D/$1 | python tools/sass_loop.py
}
show c2c_fwd.o 'Li16384E.*Li1ELi3ELb1E' "c2c 16384 (split prefetch, TWR)"
show c2c_fwd.o 'Li8192E.*Li2ELi2ELb1E' "c2c 8192 (in-place prefetch, TWR)"
show c2c_fwd.o 'Li4096E.*Li2ELi2ELb0E' "c2c 4096 (in-place prefetch)"
show c2c_fwd.o 'Li1024E.*Li3ELi1ELb1E' "c2c 1024 (double prefetch, TWR)"
show r2c.o 'Li2048E.*Li2ELi2ELb1E' "r2c 4096 (M=2048 in-place prefetch, TWR)"
show r2c_audio.o 'Li2048E.*Li2ELi2ELb1E' "stft 4096 (M=2048 in-place prefetch, TWR)"
show c2r.o 'Li2048E.*Li2ELi2ELb1E' "c2r 4096 (M=2048 in-place prefetch, TWR)"
show r2c.o 'Li16384E.*Li1ELi3ELb0E' "r2c 32768 (M=16384 split prefetch)"
show c2r.o 'Li16384E.*Li1ELi2ELb0E' "c2r 32768 (M=16384 in-place prefetch)"
