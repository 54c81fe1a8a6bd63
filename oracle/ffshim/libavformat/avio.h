/* fake-libav shim (oracle/ffshim/ffshim.h): lets the reference sources compile without FFmpeg */
#include "../ffshim.h"
