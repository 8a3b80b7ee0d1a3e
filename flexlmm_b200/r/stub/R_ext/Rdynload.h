#include "../Rinternals.h"
