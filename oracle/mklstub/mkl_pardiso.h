#include "mkl_types.h"
