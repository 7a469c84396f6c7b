#!/bin/bash
# compile one reference source against the shim (development helper)
REF=/root/reference/src
HERE=$(cd $(dirname $0) && pwd)
PYINC=$(python3 -c "import sysconfig; print(sysconfig.get_paths()['include'])")
PBINC=$(python3 -c "import pybind11; print(pybind11.get_include())")
g++ -std=c++17 -O2 -fPIC -fopenmp -w -I$HERE -I$REF -I$PYINC -I$PBINC -c $REF/limap/$1 -o /tmp/ref_$(basename $1).o 2>&1 | grep -E "error|Error" | head -${2:-40}
