#!/bin/sh
# Builds the debugging emulator (tests/emu/libemu.so). Test infrastructure only.
cd "$(dirname "$0")" && g++ -std=c++17 -O1 -g -fPIC -shared -I. -o libemu.so emu.cpp -Wno-unused-function
