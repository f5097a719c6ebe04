"""`Social`: follower / followee access over the relation list (reference: data/social.py:4-73).

The reference also materialises a SparseMatrix of the relations (`trustMatrix`, used by `row/col/elem`); the
engine-side models only need the two adjacency dictionaries and the first-appearance user order, so the matrix
views are served from the dictionaries."""
from collections import defaultdict


class Social(object):
    def __init__(self, conf, relation=None):
        self.config = conf
        self.user = {}                              # first-appearance order of the users of the relation list
        self.relation = relation
        self.followees = defaultdict(dict)
        self.followers = defaultdict(dict)
        for u1, u2, w in self.relation:
            self.followees[u1][u2] = w
            self.followers[u2][u1] = w
            for u in (u1, u2):
                if u not in self.user:
                    self.user[u] = len(self.user)

    def weight(self, u1, u2):
        return self.followees[u1][u2] if (u1 in self.followees and u2 in self.followees[u1]) else 0

    def trustSize(self):
        return (len(self.user), len(self.user), sum(len(v) for v in self.followees.values()))

    def getFollowers(self, u):
        return self.followers[u] if u in self.followers else {}

    def getFollowees(self, u):
        return self.followees[u] if u in self.followees else {}

    def hasFollowee(self, u1, u2):
        return u1 in self.followees and u2 in self.followees[u1]

    def hasFollower(self, u1, u2):
        return u1 in self.followers and u2 in self.followers[u1]
