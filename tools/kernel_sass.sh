#!/bin/bash
# usage: tools/kernel_sass.sh <libprl.so> <mangled-name substring>  -> SASS of that kernel on stdout (one instruction per line)
set -e
LIB=$(readlink -f $1); KERN=$2
TMP=$(mktemp -d); (cd $TMP && cuobjdump -xelf all $LIB > /dev/null)
nvdisasm -c $TMP/prl_api.sm_100a.cubin | awk -v k="$KERN" '/^\.text\./{f=index($0,k)>0} f&&/^[ \t]+\/\*[0-9a-f]+\*\//{print}'
rm -rf $TMP
