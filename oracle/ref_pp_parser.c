/* oracle/ref_pp_parser.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Hand-written precedence-climbing parser providing PPparse(), the entry point
 * the reference generates with bison from src/frontend/parse-bison.y (bison is
 * not available in this image).  It accepts the same language and builds the
 * same pnode trees through the reference's own PP_mk*node constructors
 * (src/frontend/parse.c), so `let`, `print`, `meas`, `alter` expressions work in
 * the oracle binary.  Operator table follows parse-bison.y:83-95.
 */
#include "ngspice/ngspice.h"
#include "ngspice/fteparse.h"
#include <stdio.h>
#include <stdlib.h>
#include "parse.h"
#include "parse-bison.h"
#include "parse-bison-y.h"

typedef struct {
    char **line;
    int tok;
    YYSTYPE val;
    struct PPltype loc;
    const char *last_stop;
    int failed;
} pp_t;

static void pp_next(pp_t *p)
{
    p->last_stop = p->loc.stop;
    p->val.num = 0.0;
    p->tok = PPlex(&p->val, &p->loc, p->line);
}

/* binary-operator precedence levels (higher binds tighter); 0 = not binary */
enum { L_TERN = 1, L_OR, L_AND, L_CMP, L_NOT, L_COMMA, L_ADD, L_MUL, L_NEG, L_POW, L_IDX };

static int pp_binprec(int tok, int *right, int *op)
{
    *right = 0;
    switch (tok) {
    case '|': *op = PT_OP_OR;  return L_OR;
    case '&': *op = PT_OP_AND; return L_AND;
    case '=': *op = PT_OP_EQ;  return L_CMP;
    case TOK_NE: *op = PT_OP_NE; return L_CMP;
    case TOK_LE: *op = PT_OP_LE; return L_CMP;
    case TOK_GE: *op = PT_OP_GE; return L_CMP;
    case '<': *op = PT_OP_LT;  return L_CMP;
    case '>': *op = PT_OP_GT;  return L_CMP;
    case ',': *op = PT_OP_COMMA; *right = 1; return L_COMMA;
    case '+': *op = PT_OP_PLUS;  return L_ADD;
    case '-': *op = PT_OP_MINUS; return L_ADD;
    case '*': *op = PT_OP_TIMES; return L_MUL;
    case '/': *op = PT_OP_DIVIDE; return L_MUL;
    case '%': *op = PT_OP_MOD;   return L_MUL;
    case '^': *op = PT_OP_POWER; *right = 1; return L_POW;
    default: return 0;
    }
}

static struct pnode *pp_exp(pp_t *p, int minprec);

static struct pnode *pp_primary(pp_t *p)
{
    struct pnode *n;
    if (p->tok == TOK_NUM) {
        n = PP_mknnode(p->val.num);
        pp_next(p);
        return n;
    }
    if (p->tok == TOK_STR) {
        const char *s = p->val.str;
        pp_next(p);
        if (p->tok == '(') {            /* function application is favoured */
            struct pnode *arg;
            pp_next(p);
            arg = pp_exp(p, L_TERN);
            if (!arg || p->tok != ')') { p->failed = 1; txfree(s); return NULL; }
            pp_next(p);
            n = PP_mkfnode(s, arg);
            txfree(s);
            if (!n) p->failed = 1;
            return n;
        }
        n = PP_mksnode(s);
        txfree(s);
        return n;
    }
    if (p->tok == '(') {
        pp_next(p);
        n = pp_exp(p, L_TERN);
        if (!n || p->tok != ')') { p->failed = 1; return NULL; }
        pp_next(p);
        return n;
    }
    if (p->tok == '-') {
        pp_next(p);
        n = pp_exp(p, L_NEG + 1);
        if (!n) return NULL;
        return PP_mkunode(PT_OP_UMINUS, n);
    }
    if (p->tok == '~') {
        pp_next(p);
        n = pp_exp(p, L_NOT + 1);
        if (!n) return NULL;
        return PP_mkunode(PT_OP_NOT, n);
    }
    p->failed = 1;
    return NULL;
}

static struct pnode *pp_exp(pp_t *p, int minprec)
{
    struct pnode *lhs = pp_primary(p);
    if (!lhs) return NULL;
    for (;;) {
        int right, op, prec;
        if (p->tok == '[' && L_IDX >= minprec) {
            struct pnode *idx;
            pp_next(p);
            if (p->tok == '[') {
                pp_next(p);
                idx = pp_exp(p, L_TERN);
                if (!idx || p->tok != ']') { p->failed = 1; return NULL; }
                pp_next(p);
                if (p->tok != ']') { p->failed = 1; return NULL; }
                pp_next(p);
                lhs = PP_mkbnode(PT_OP_RANGE, lhs, idx);
            } else {
                idx = pp_exp(p, L_TERN);
                if (!idx || p->tok != ']') { p->failed = 1; return NULL; }
                pp_next(p);
                lhs = PP_mkbnode(PT_OP_INDX, lhs, idx);
            }
            continue;
        }
        if (p->tok == '?' && L_TERN >= minprec) {
            struct pnode *a, *b;
            pp_next(p);
            a = pp_exp(p, L_TERN);
            if (!a || p->tok != ':') { p->failed = 1; return NULL; }
            pp_next(p);
            b = pp_exp(p, L_TERN);
            if (!b) return NULL;
            lhs = PP_mkbnode(PT_OP_TERNARY, lhs, PP_mkbnode(PT_OP_COMMA, a, b));
            continue;
        }
        prec = pp_binprec(p->tok, &right, &op);
        if (!prec || prec < minprec) break;
        pp_next(p);
        {
            struct pnode *rhs = pp_exp(p, right ? prec : prec + 1);
            if (!rhs) return NULL;
            lhs = PP_mkbnode(op, lhs, rhs);
        }
    }
    return lhs;
}

int PPparse(char **line, struct pnode **retval)
{
    pp_t P;
    struct pnode *head = NULL, *tail = NULL;
    char *keep = *line;
    P.line = line; P.failed = 0; P.loc.start = P.loc.stop = NULL; P.last_stop = NULL;
    pp_next(&P);
    *retval = NULL;
    while (P.tok != 0) {
        const char *start = P.loc.start;
        struct pnode *e = pp_exp(&P, L_TERN);
        if (!e || P.failed) {
            fprintf(stderr, "PPparse: syntax error in line segment\n   %s\n", keep);
            return 1;
        }
        e->pn_name = copy_substring(start, P.last_stop);
        if (!head) head = e; else { tail->pn_next = e; e->pn_use++; }
        tail = e;
    }
    *retval = head;
    return 0;
}
