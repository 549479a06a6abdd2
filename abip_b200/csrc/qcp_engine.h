// qcp_engine.h -- internal header of the QCP engine.
#pragma once
#include "../../include/abip_gpu.h"
