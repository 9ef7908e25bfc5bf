#!/bin/bash
# A/B builds of libpolus_b200.so with the same ABI:  tools/build_variant.sh NAME "-DFLAG=.. ..." file.cu [file.cu ...]
# compiles the named sources with the extra flags into polus_b200/csrc/build_NAME/, links them with the default build's
# other objects into polus_b200/libpolus_b200_NAME.so; select at run time with POLUS_LIB=polus_b200/libpolus_b200_NAME.so.
set -e
NAME=$1; FLAGS=$2; shift 2
cd "$(dirname "$0")/../polus_b200/csrc"
make -s -j 8 > /dev/null
mkdir -p build_$NAME
NCCL_HOME=$(python3 -c "import nvidia, os; print(os.path.join(list(nvidia.__path__)[0], 'nccl'))")
OBJS=""
for o in build/*.o; do
  b=$(basename $o .o)
  if [[ " $* " == *" $b.cu "* ]]; then
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v \
      $FLAGS -I$NCCL_HOME/include -c $b.cu -o build_$NAME/$b.o 2> build_$NAME/$b.ptxas.log || { tail -20 build_$NAME/$b.ptxas.log; exit 1; }
    OBJS="$OBJS build_$NAME/$b.o"
  else
    OBJS="$OBJS $o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpolus_b200_$NAME.so $OBJS -L$NCCL_HOME/lib -l:libnccl.so.2 -Xlinker -rpath -Xlinker $NCCL_HOME/lib -lcudart
echo "built polus_b200/libpolus_b200_$NAME.so"
