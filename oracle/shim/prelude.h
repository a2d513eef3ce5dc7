/* oracle/shim/prelude.h -- force-included (-include) ahead of every reference TU.
 * TEST INFRASTRUCTURE ONLY. Pulls in the std headers first, then maps the MSVC-only
 * `std::exception(const char*)` constructor used by the reference
 * (src/decoder/BrotligHuffmanTable.cpp:203) onto std::runtime_error. */
#pragma once
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <exception>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <variant>
#include <vector>
#define exception runtime_error
