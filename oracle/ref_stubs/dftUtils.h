// STUB (test infrastructure): the reference's include/dftUtils.h declares deal.II-typed helpers none of the
// translation units compiled into oracle/_ref call.
#pragma once
#include <headers.h>
