/* admm_oracle.c -- CPU restatement of the per-iteration ADMM algebra of
 * MCONTACT::CONTACT_ANALYSIS.  TEST INFRASTRUCTURE ONLY (see mgpis_oracle.c).
 * (filled in with the ADMM rows of SURVEY.md §8a) */
int orc_admm_placeholder(void) { return 0; }
