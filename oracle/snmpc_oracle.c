/* SNMPC oracle: filled in later */
#include "nmpc_oracle.h"

