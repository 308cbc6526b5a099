/* oracle/ref_pt_parser.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Hand-written precedence-climbing parser providing PTparse(), which the
 * reference generates with bison from src/spicelib/parser/inpptree-parser.y
 * (bison is not available here).  Same language, same trees, built through the
 * reference's PT_mk*node constructors (src/spicelib/parser/inpptree.c).
 * Precedence table: inpptree-parser.y:58-69.
 */
#include "ngspice/ngspice.h"
#include "ngspice/inpptree.h"
#include "ngspice/inpdefs.h"
#include <stdio.h>
#include <stdlib.h>
#include "inpptree-parser.h"
#include "inpptree-parser-y.h"

typedef struct {
    char **line;
    CKTcircuit *ckt;
    int tok;
    YYSTYPE val;
    struct PTltype loc;
    char *last_stop;
    int failed;
} pt_t;

static void pt_next(pt_t *p)
{
    p->last_stop = p->loc.stop;
    p->val.num = 0.0;
    p->tok = PTlex(&p->val, &p->loc, p->line);
}

enum { Q_TERN = 1, Q_OR, Q_AND, Q_EQ, Q_REL, Q_ADD, Q_MUL, Q_NEG, Q_POW };

static INPparseNode *pt_exp(pt_t *p, int minprec);

static INPparseNode *pt_primary(pt_t *p)
{
    INPparseNode *n;
    switch (p->tok) {
    case TOK_NUM:
        n = PT_mknnode(p->val.num); pt_next(p); return n;
    case TOK_pnode:
        n = p->val.pnode; pt_next(p); return n;
    case TOK_STR: {
        const char *s = p->val.str;
        pt_next(p);
        if (p->tok == '(') {
            INPparseNode *args;
            pt_next(p);
            args = pt_exp(p, Q_TERN);
            if (!args) { txfree(s); return NULL; }
            while (p->tok == ',') {
                INPparseNode *a;
                pt_next(p);
                a = pt_exp(p, Q_TERN);
                if (!a) { txfree(s); return NULL; }
                args = PT_mkbnode(",", args, a);
            }
            if (p->tok != ')') { p->failed = 1; txfree(s); return NULL; }
            pt_next(p);
            n = PT_mkfnode(s, args);
            txfree(s);
            if (!n) p->failed = 1;
            return n;
        }
        n = PT_mksnode(s, p->ckt);
        txfree(s);
        return n;
    }
    case '(':
        pt_next(p);
        n = pt_exp(p, Q_TERN);
        if (!n || p->tok != ')') { p->failed = 1; return NULL; }
        pt_next(p);
        return n;
    case '-':
        pt_next(p);
        n = pt_exp(p, Q_NEG + 1);
        return n ? PT_mkfnode("-", n) : NULL;
    case '+':
        pt_next(p);
        return pt_exp(p, Q_NEG + 1);
    case '!':
        pt_next(p);
        n = pt_exp(p, Q_NEG + 1);
        return n ? PT_mkfnode("eq0", n) : NULL;
    default:
        p->failed = 1;
        return NULL;
    }
}

static INPparseNode *pt_exp(pt_t *p, int minprec)
{
    INPparseNode *lhs = pt_primary(p);
    if (!lhs) return NULL;
    for (;;) {
        int prec = 0;
        const char *bop = NULL, *cmp = NULL;
        int logic = 0;
        switch (p->tok) {
        case '?': prec = Q_TERN; break;
        case TOK_OR:  prec = Q_OR;  logic = 1; break;
        case TOK_AND: prec = Q_AND; logic = 2; break;
        case TOK_EQ: prec = Q_EQ; cmp = "eq0"; break;
        case TOK_NE: prec = Q_EQ; cmp = "ne0"; break;
        case TOK_GT: prec = Q_REL; cmp = "gt0"; break;
        case TOK_LT: prec = Q_REL; cmp = "lt0"; break;
        case TOK_GE: prec = Q_REL; cmp = "ge0"; break;
        case TOK_LE: prec = Q_REL; cmp = "le0"; break;
        case '+': prec = Q_ADD; bop = "+"; break;
        case '-': prec = Q_ADD; bop = "-"; break;
        case '*': prec = Q_MUL; bop = "*"; break;
        case '/': prec = Q_MUL; bop = "/"; break;
        case '^': prec = Q_POW; bop = "^"; break;
        default: break;
        }
        if (!prec || prec < minprec) break;
        if (p->tok == '?') {
            INPparseNode *a, *b;
            pt_next(p);
            a = pt_exp(p, Q_TERN);
            if (!a || p->tok != ':') { p->failed = 1; return NULL; }
            pt_next(p);
            b = pt_exp(p, Q_TERN);
            if (!b) return NULL;
            lhs = PT_mkfnode("ternary_fcn", PT_mkbnode(",", PT_mkbnode(",", lhs, a), b));
            continue;
        }
        pt_next(p);
        {
            INPparseNode *rhs = pt_exp(p, prec + 1);   /* all binary ops are %left */
            if (!rhs) return NULL;
            if (bop)
                lhs = PT_mkbnode(bop, lhs, rhs);
            else if (cmp)
                lhs = PT_mkfnode(cmp, PT_mkbnode("-", lhs, rhs));
            else if (logic == 1)
                lhs = PT_mkfnode("ne0", PT_mkbnode("+", PT_mkfnode("ne0", lhs), PT_mkfnode("ne0", rhs)));
            else
                lhs = PT_mkfnode("eq0", PT_mkbnode("+", PT_mkfnode("eq0", lhs), PT_mkfnode("eq0", rhs)));
        }
    }
    return lhs;
}

int PTparse(char **line, INPparseNode **retval, CKTcircuit *ckt)
{
    pt_t P;
    INPparseNode *e;
    P.line = line; P.ckt = ckt; P.failed = 0; P.loc.start = P.loc.stop = NULL; P.last_stop = NULL;
    pt_next(&P);
    e = pt_exp(&P, Q_TERN);
    if (!e || P.failed) {
        fprintf(stderr, "PTparse: syntax error in expression near %s\n", *line ? *line : "");
        controlled_exit(EXIT_BAD);
    }
    *retval = e;
    *line = P.last_stop;      /* stop after the last token that belongs to the expression */
    return 0;
}
