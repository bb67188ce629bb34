#define MRB_TEAM_SCN MRB_SIMPLE
#define MRB_TEAM_TAG simple
#define MRB_TEAM_N 16
#include "kern_team.inc.h"
