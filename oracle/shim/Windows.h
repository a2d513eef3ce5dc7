/* oracle/shim: empty stand-in so the reference's BrotligCommon.h (#include <Windows.h>) compiles on Linux. */
