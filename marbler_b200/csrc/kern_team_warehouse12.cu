#define MRB_TEAM_SCN MRB_WAREHOUSE
#define MRB_TEAM_TAG warehouse
#define MRB_TEAM_N 12
#include "kern_team.inc.h"
