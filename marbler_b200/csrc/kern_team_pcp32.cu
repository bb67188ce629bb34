#define MRB_TEAM_SCN MRB_PCP
#define MRB_TEAM_TAG pcp
#define MRB_TEAM_N 32
#define MRB_TEAM_WITH_QP
#include "kern_team.inc.h"
