#define GB_TAG g2
#define GB_NW 2
#define GB_KM 1
#include "gb_inst.inc"
