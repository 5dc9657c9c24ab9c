// TEST INFRASTRUCTURE ONLY.  Prelude of oracle/_ref/compute_history.gen.cpp (see oracle/Makefile): the
// includes and `using` that gcn/history.cpp:1-8 provide for the commented-out `compute_history`
// (gcn/history.cpp:10-37), plus a null stream that swallows the GFLOP/s line it prints on every call.
#pragma once
#include <omp.h>

#include <chrono>
#include <iostream>
#include <vector>
using namespace std;
using namespace std::chrono;
static std::ostream sgcn_null_out(nullptr);
#define cout sgcn_null_out
