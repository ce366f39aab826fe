// ekgsim_b200/host/compat/refglue_shim.h -- forced include (g++ -include) that lets the REFERENCE's
// own evaluation glue (sim.cpp, main.cpp; compiled from the reference tree, never copied) build against
// the B200 facade instead of the reference's simlib.  It pre-empts the include guards of the four
// hot-path headers (simlib/sim_lib.h, simulator.h, matrix.h, Wohlfart.h) and of simlib/support.h, pulls
// the reference's non-hot-path helpers that its glue expects to arrive through them (Ini.h,
// columnFile.h), and then provides the same names from ekgsim_b200/host/sim_lib.h.
// Used by the test harness that builds the reference glue on top of this facade.
#pragma once
#define SIM_LIB_H_INCLUDED
#define SIMULATOR_H_INCLUDED
#define MATRIX_H_INCLUDED
#define WOHLFART_H_INCLUDED
#define SUPPORT_H_INCLUDED

#include <algorithm>
#include <cassert>
#include <cmath>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <queue>
#include <sstream>
#include <string>
#include <vector>

#include <Ini.h>         // reference: copyOfLibs/Ini.h
#include "columnFile.h"  // reference: simlib/columnFile.h

#include "ekgsim_b200/host/sim_lib.h"

using SimLib::WohlfartPlus;

template <class T>
std::basic_string<T> filenameExtension(const std::basic_string<T>& filename) {
	for (int i = (int)filename.size() - 1; i >= 0; --i)
		if (filename[i] == '.') return std::basic_string<T>(filename.begin() + i, filename.end());
	return std::basic_string<T>();
}
