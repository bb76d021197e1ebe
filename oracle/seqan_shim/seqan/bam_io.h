// SeqAn 1.4 stand-in header (test infrastructure; see shim_impl.h).
#pragma once
#include "shim_impl.h"
