#!/bin/sh
# Turns the device headers that NVRTC compiles at run time into C byte arrays (jit_runtime.cu includes the result).
# usage: embed_sources.sh <include-name>=<file> ...
i=0
names=""
for spec in "$@"; do
    name="${spec%%=*}"; file="${spec#*=}"
    echo "static const unsigned char kEmbedded_$i[] = {"
    od -An -v -tx1 "$file" | sed 's/ \([0-9a-f][0-9a-f]\)/0x\1,/g'
    echo "0x00};"
    names="$names    {\"$name\", kEmbedded_$i},\n"
    i=$((i + 1))
done
printf "static const EmbeddedSource kEmbeddedSources[] = {\n$names    {nullptr, nullptr}};\n"
