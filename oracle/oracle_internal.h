/* Shared between the oracle's translation units (TEST INFRASTRUCTURE, NOT PRODUCT CODE). */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H
extern int oracle_g_warmup_steps;
extern double oracle_g_loop_seconds;
double oracle_now_seconds(void);
#endif
