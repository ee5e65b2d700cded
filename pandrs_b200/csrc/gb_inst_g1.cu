#define GB_TAG g1
#define GB_NW 1
#define GB_KM 1
#include "gb_inst.inc"
