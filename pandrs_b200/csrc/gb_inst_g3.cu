#define GB_TAG g3
#define GB_NW 3
#define GB_KM 1
#include "gb_inst.inc"
