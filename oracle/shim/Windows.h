/* settingsFile.h:2 spells it with a capital W; same shim. */
#include "windows.h"
