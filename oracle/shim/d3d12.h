/* oracle/shim: empty stand-in for <d3d12.h> (BrotligCommon.h includes it; the decode path uses nothing from it). */
