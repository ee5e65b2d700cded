#define GB_TAG k1
#define GB_NW 1
#define GB_KM 0
#include "gb_inst.inc"
